// pm_types.h -- data shared by the host-side problem compiler (pm_build.cu) and the platoon-MIQP
// kernel (pm_kernel.cu): the "compiled structure" of one MPC formulation.
//
// A formulation (centralized cent_mld.py:48-177, event-based fleet_event_based.py:72-291, naive ADMM
// fleet_naive_admm.py:63-237, g-ADMM fleet_g_admm.py:55-158, local fleet_decent_mld.py:61-208) is
// compiled ONCE -- like the reference builds its Gurobi model once -- into dense matrices over the
// decision vector
//     z = [ v_{i,k}  (i = 0..nl-1 vehicles, k = 1..N) ,  extras (free copies) ]
// and the per-solve parameter vector
//     pvec = [ x0 (2 nl) , params (npar) , 1 ].
// Everything that differs between the problems of a batch is pvec, the masses and the result.
#pragma once
#include <stdint.h>

namespace hvp {

constexpr int PM_MAXMODES = 12;

struct PmModel {                 // PWA modes of one vehicle: v+ = a v + b u + c for lo <= v <= hi
    int R;                       // number of modes (7: pwa_gear, 12: friction x gear)
    double cf[PM_MAXMODES];      // a = 1 - cf/m
    double bg[PM_MAXMODES];      // b = bg/m
    double dd[PM_MAXMODES];      // c = -mu g - dd/m
    double lo[PM_MAXMODES], hi[PM_MAXMODES];
    int gear[PM_MAXMODES];       // gear (1..6) reported for the mode
    double mug;
};

struct PmDev {                   // kernel argument
    int nl, N, ne, nv, ld;       // vehicles, horizon, extras, variables, padded (odd) leading dimension
    int npar, npv;               // params per problem; npv = 2 nl + npar + 1
    int nres, nlin, ng, n0;      // residuals, param-linear cost terms, generic rows, constant rows
    int depth;                   // nl * N branching decisions (stage-major: d = k nl + i)
    int max_nodes;
    // 1-norm cost (quadratic_cost = False, cent_mld.py:58-61): every cost term w |e| is a PAIR of L1-penalised rows
    // (e <= 0 and -e <= 0, weight w), the input cost qu |u| of a fixed stage a pair of on-the-fly rows, and a node LP is
    // solved by the proximal-point method on the same bounded-multiplier dual active-set code: H = rho_px I around a
    // centre z_k that moves to the solution until it stops moving (finite for an LP; one or two rounds in practice).
    int one_norm;
    double rho_px;
    double ppa_stall;            // a proximal round whose LP objective decreases by less than this (relative) ends the node
    int sibling;                 // prune siblings by the dual of their solved parent (pm_kernel.cu parent_info)
    double mip_gap;              // relative pruning gap (0: proven optimal)
    long long time_limit_ns;     // per-problem budget on the device clock (0: none)
    PmModel M;
    double qu, w, vmin, vmax, pmin, pmax, umin, umax, a_acc, a_dec, tight;
    // device arrays (shared by all problems; read-only, L1/L2 resident)
    const double* H0;            // [nv][nv]      2 R'WR
    const double* H0inv;         // [nv][nv]
    const double* Cres;          // [nres][npv]   residual constants c_r = Cres[r] . pvec
    const double* wres;          // [nres]
    const double* RW2;           // [nv][nres]    2 w_r R[r][j]
    const double* La;            // [nlin][npv]   cost += (La.pvec) * (Lz.z + Lp.pvec)
    const double* Lz;            // [nlin][nv]
    const double* Lp;            // [nlin][npv]
    const double* AT;            // [nv][ng]      generic rows, column-major: row r is sum_j AT[j][r] z_j <= BR[r].pvec
    const double* BR;            // [ng][npv]
    const double* wmax;          // [ng]          L1 penalty weight (soft) or +inf (hard)
    const double* B0;            // [n0][npv]     constant rows: B0[r].pvec <= 0
    const double* w0;            // [n0]
    // shared-structure multi-RHS product (tensor-core precompute): Y[b] = W . pvec_b with
    // W = [GC ; BR ; B0 ; Cres ; La ; Lp] ([mw][kw], zero padded to multiples of 8 x 4):
    //   GC = RW2.Cres + Lz'.La  -> g0 = GC.pvec      BR -> generic right-hand sides      B0 -> constant rows
    //   Cres, La, Lp -> residual constants c_r and the factors of the param-linear terms:
    //                   c0 = sum_r w_r c_r^2 + sum_l a_l p_l
    const double* W;
    int mw, kw;
    // Free copies that appear in NO row (the velocity copies of the naive-ADMM formulation) are eliminated when the
    // formulation is compiled: with z = (zK, zE) and H0 = [[A, B], [B', C]], minimising over zE gives the reduced
    // Hessian A - B C^-1 B', gradient gK - B C^-1 gE and constant c - gE' C^-1 gE / 2, and zE* = -C^-1 (B' zK + gE).
    // nv counts the KEPT variables; ne stays the number of extras REPORTED.  The precompute appends gE and
    // hE = -C^-1 gE to Y (rows o_yel .. o_yel + 2 nel); emap[e] >= 0: kept extra (index among the kept extras),
    // < 0: eliminated extra -1 - emap[e]; Rz = -C^-1 B' ([nel][nv]).
    int nel, o_yel;
    const int* emap;             // [ne]
    const double* Rz;            // [nel][nv]
    // shared-memory carve-up (offsets in doubles / ints), filled by pm_layout()
    int o_hinv, o_ginv, o_nact, o_vec, o_cres, o_bgen, o_pvec, o_misc, smem_doubles;
    int o_int, smem_bytes;
};

// Tree splitting of heavy problems (pm_kernel.cu): pass 1 solves every problem under a node budget and
// appends the ones that exceed it to `flagged`; pass 2 gives each flagged problem `M` warps, which share the
// work by the ordinal of the depth-`D` mode prefix (ordinal mod M) and exchange the incumbent through
// `inc_shared` (atomicMin on an order-preserving key); pass 3 keeps the best sub-result.
struct PmSplit {
    int mode;                          // 0 plain, 1 budgeted pass, 2 sub-tree pass, 3 sharded pass (one rank's share)
    int budget, cap, M, D;
    int rank, world;                   // mode 3: this device takes the prefix ordinals o with o % world == rank
    int* nflag;                        // [1] flagged problems so far
    int* flagged;                      // [cap] their batch indices
    unsigned long long* inc_shared;    // [cap] best objective known for each (order-preserving key)
    // Adoption in the sub-tree pass (mode 2): a worker that finds the work counter exhausted WAITS; a busy worker that
    // sees waiting ones gives away the shallowest untried sibling of its depth-first stack as a mode prefix.
    int ad;                            // 1: on in this launch
    int ad_free;                       // a donated sub-tree keeps at least this many free levels
    int pool_cap, mail_cap, mail_stride;
    int pool_base;                     // first result slot of the pool (mode 2: cap * M, mode 3: batch * M)
    unsigned long long* ad_count;      // started workers << 32 | waiting workers
    int* pool_used;                    // result slots handed to adopted sub-trees so far (slot p lives at cap * M + p)
    int* pool_owner;                   // [pool_cap] flagged index of the problem a slot belongs to
    int* mail_state;                   // [mail_cap] 0 busy, 1 waiting, 3 claimed by a donor, 2 job posted
    int* mail_job;                     // [mail_cap][mail_stride] flagged index, prefix length, slot, prefix modes
};

// device scratch of the sub-tree pass: outputs of cap * M work items, same layout as the real outputs
struct PmScratch {
    PmSplit sp;
    double *u, *x, *extra, *obj;
    int32_t *modes, *status, *nodes, *iters;
};

}  // namespace hvp
