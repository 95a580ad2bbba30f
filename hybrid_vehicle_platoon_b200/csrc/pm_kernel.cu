// pm_kernel.cu -- platoon MIQP kernel: ONE WARP per mixed-integer QP, dense node QPs in shared memory.
//
// Replaces the Gurobi solves behind MpcMldCent.solve_mpc (mpcs/cent_mld.py:9-182 on dmpcpwa
// MpcMldCentDecup), the event-based LocalMpc (fleet_event_based.py:26-340), the naive-ADMM
// LocalMpcADMM (fleet_naive_admm.py:24-258), the g-ADMM fixed-mode LocalMpc (fleet_g_admm.py:22-205)
// and the MpcGear variants (mpcs/mpc_gear.py:8-170) -- every formulation whose decision vector is
// larger than the per-vehicle problem of local_miqp.cu.  The formulation itself (cost residuals,
// coupling rows) arrives compiled into dense matrices (pm_types.h / pm_build.cu); this file is the
// solver:
//
//   * decision vector z = [future velocities v_{i,k} of the nl local vehicles ; free copies];
//     positions are prefix sums, inputs are the affine map u = (v+ - a v - c)/b of the active mode;
//   * branch and bound, depth first, over the (stage, vehicle) mode decisions in stage-major order;
//     a node fixes decisions 0..L-1 and relaxes the rest (no input cost, no input limits, no region
//     rows: a valid lower bound); children ordered by distance to the relaxed velocity, pruned by
//     interval reachability and by the incumbent (gap 0);
//   * node QP: dual active-set (Goldfarb-Idnani) with bounded multipliers for the L1-penalised
//     (slack-eliminated) rows.  The warp holds H^-1 (nv x nv), the dense normals of the active rows
//     and the inverse of N'H^-1N in shared memory; lane j owns variable j / slot j, mat-vecs are one
//     row per lane, reductions and arg-min/max are warp shuffles.  H^-1 follows the search by
//     Sherman-Morrison rank-1 up/down-dates (one per fixed decision), re-seeded from the shared
//     H0^-1 whenever the search returns to the root.
//
// HBM traffic per problem is parameters in + solution out; the shared structure matrices are read
// through L1/L2.  All arithmetic is FP64.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "hvp_internal.h"
#include "miqp_core.cuh"   // status codes, hvp_now_ns, hvp_cut
#include "pm_types.h"

namespace hvp {

namespace {

constexpr unsigned FULL = 0xffffffffu;
// PT_USP / PT_USN: the pair of soft rows  b u <= 0 / -b u <= 0  (weight qu / b) of a fixed stage under the 1-norm cost
enum : int { PT_UB = 0, PT_LB, PT_UHI, PT_ULO, PT_ACC, PT_DEC, PT_PHI, PT_PLO, PT_USP, PT_USN, PT_GEN };
enum : int { PS_NEXT = 0, PS_BUILD, PS_SELECT, PS_STEP, PS_DONE };
#define PM_ID(t, idx) ((t) * 4096 + (idx))
#define LANES(j, n) _Pragma("unroll 1") for (int j = lane; j < (n); j += GW)

// reciprocal to ~1 ulp without the slow paths / code size of the IEEE division sequence
__device__ __forceinline__ double rcp(double v) {
    // MUFU.RCP64H seed (20 bits, one instruction, no slow path) + two Newton steps (40, 80 bits)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
    r = r * (2.0 - v * r);
    r = r * (2.0 - v * r);
    return r;
}
// dot product with two independent accumulation chains (the solver is latency bound);
// `sa` = stride of a in doubles
__device__ __forceinline__ double dot2(const double* a, int sa, const double* b, int n) {
    // four independent chains, 32-bit index arithmetic: the solver is latency bound (r01l ncu: this function is a
    // quarter of all stall samples, almost all of them dependency waits)
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int c = 0;
    const double* ap = a;
#pragma unroll 1
    for (; c + 3 < n; c += 4, ap += 4 * sa, b += 4) {
        s0 += ap[0] * b[0];
        s1 += ap[sa] * b[1];
        s2 += ap[2 * sa] * b[2];
        s3 += ap[3 * sa] * b[3];
    }
#pragma unroll 1
    for (; c < n; ++c, ap += sa, ++b) s0 += ap[0] * b[0];
    return (s0 + s1) + (s2 + s3);
}
// GW = lanes of the group that owns one problem (32, or 16 so that a warp carries two small problems);
// gm = the group's lane mask.  Shuffles use width GW, so lane indices are relative to the group.
template <int GW>
__device__ __forceinline__ double wsum(unsigned gm, double v) {
#pragma unroll
    for (int o = GW / 2; o; o >>= 1) v += __shfl_xor_sync(gm, v, o, GW);
    return v;
}
// arg-max over the warp with the smaller id winning ties (every lane gets the same answer)
template <int GW>
__device__ __forceinline__ void wargmax(unsigned gm, double& v, int& id) {
#pragma unroll
    for (int o = GW / 2; o; o >>= 1) {
        const double ov = __shfl_xor_sync(gm, v, o, GW);
        const int oid = __shfl_xor_sync(gm, id, o, GW);
        if (ov > v || (ov == v && oid < id)) { v = ov; id = oid; }
    }
}
template <int GW>
__device__ __forceinline__ void wargmin(unsigned gm, double& v, int& id) {
#pragma unroll
    for (int o = GW / 2; o; o >>= 1) {
        const double ov = __shfl_xor_sync(gm, v, o, GW);
        const int oid = __shfl_xor_sync(gm, id, o, GW);
        if (ov < v || (ov == v && oid < id)) { v = ov; id = oid; }
    }
}

// order-preserving map double -> uint64 (so that atomicMin orders objectives)
__device__ __forceinline__ unsigned long long dkey(double d) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dval(unsigned long long k) {
    const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

// ONE: the 1-norm (MILP) variant.  A template parameter, not a run-time flag: this kernel is bound by instruction fetch, and the
// proximal-point code of the LP nodes, never executed by a 2-norm launch, still cost it 7 % (r02 measurement).
// FX: the fixed-sequence QP of g-ADMM (fixed_modes given): no search at all, so the branch-and-bound code is compiled out.
// AD: adoption of sub-trees between the workers of the sub-tree pass (PmSplit::ad) -- its own instantiation, launched for
// that pass only, because even never-executed code costs this kernel time (dealing prefixes from a counter: -23 % on
// every launch when it was a run-time path).
template <int GW, bool ONE, bool FX = false, bool AD = false>
struct Warp {
    const PmDev& S;
    int lane;
    unsigned gm;                       // lane mask of this group
    // shared-memory views
    double *Hinv, *Ginv, *Nact;
    double *x, *g0, *gn, *yp, *wv, *zd, *dv, *rv, *lam, *np_, *best, *zc;     // zd: right-hand sides of the active rows (1-norm)
    double bp;
    double *cres, *bgen, *pvec;
    double *inv_m, *pc, *v0, *xstar, *rlo, *rhi, *am, *bm, *cm, *amax, *amin;
    // Sibling bounds (levels 0..D): what the SOLVED parent of a level says about forcing its branching velocity out of the
    // relaxed value -- pf the parent's optimum (-inf: the level was opened without solving its parent), pnz the curvature
    // of its dual function along the bound row +-e_j, ptp / ptm how far that dual step may go before an active multiplier
    // leaves [0, w].  A child that fixes the velocity into a region at distance bd costs at least
    // pf + t bd - t^2 pnz / 2, t = min(bd / pnz, tmax): cheaper than building the child to find out (flat_core.cuh ALG2;
    // there it loses to SIMT divergence, here control flow is warp-uniform and fewer nodes are simply less time).
    double *pf, *pnz, *ptp, *ptm;
    int *act, *cand, *modes, *bmodes, *built, *orient, *aflag, *agen, *uor;
    double f_prev, xu_scale;           // xu_scale: size of the unconstrained minimiser of the current proximal round
    int ppa, retry;                    // proximal-point rounds / restarts done on the current node (1-norm cost)
    // warp-uniform scalars (replicated in registers)
    int state, lev, L, built_L, q, it, iters, nodes, pid, fixed;
    double inc, c0, cp, nHn, lam_p, dual;
    bool p_soft, trouble, limit, dive, timeout;
    long long t_start;
    // tree splitting (PmSplit)
    int sub_M, sub_D, sub_code, sub_ord, budget, stop_nodes;
    int job_P, fidx, ad_nwork, ad_probe;      // adoption: forced prefix length of an adopted sub-tree, flagged index, workers, probe cursor
    double own;                        // objective of this warp's own best leaf
    const double* yel;                 // gE, hE of the eliminated copies of this problem (PmDev::nel)
    unsigned long long* shared;        // incumbent shared by the warps working on one problem
    const PmSplit* sp;
    int64_t prob;

    __device__ Warp(const PmDev& S_, double* base, int lane_, unsigned gm_) : S(S_), lane(lane_), gm(gm_) {
        const int nv = S.nv, ld = S.ld;
        Hinv = base + S.o_hinv; Ginv = base + S.o_ginv; Nact = base + S.o_nact;
        double* v = base + S.o_vec;
        x = v; g0 = v + nv; gn = v + 2 * nv; yp = v + 3 * nv; wv = v + 4 * nv; zd = v + 5 * nv;
        dv = v + 6 * nv; rv = v + 7 * nv; lam = v + 8 * nv; np_ = v + 9 * nv; best = v + 10 * nv; zc = v + 11 * nv;
        cres = base + S.o_cres; bgen = base + S.o_bgen; pvec = base + S.o_pvec;
        double* m = base + S.o_misc;
        const int nl = S.nl, N = S.N, D = S.depth;
        inv_m = m; pc = m + nl; v0 = m + 2 * nl; xstar = m + 3 * nl; rlo = xstar + D + 1;
        rhi = rlo + nl * (N + 1); am = rhi + nl * (N + 1); bm = am + D; cm = bm + D;
        amax = cm + D; amin = amax + D;
        pf = amin + D; pnz = pf + (D + 1); ptp = pnz + (D + 1); ptm = ptp + (D + 1);
        int* ib = reinterpret_cast<int*>(base + S.smem_doubles);
        act = ib; cand = ib + nv; modes = cand + D + 1; bmodes = modes + D; built = bmodes + D;
        orient = built + D; aflag = orient + S.ng; agen = aflag + nv; uor = agen + S.ng;
        ppa = 0; retry = 0;
        (void)ld;
        sub_M = sub_D = sub_code = sub_ord = budget = stop_nodes = 0; job_P = fidx = ad_nwork = ad_probe = 0; shared = nullptr; sp = nullptr; prob = 0; own = HUGE_VAL;
    }

    __device__ __forceinline__ double ma(int i, int r) const { return 1.0 - S.M.cf[r] * inv_m[i]; }
    __device__ __forceinline__ double mb(int i, int r) const { return S.M.bg[r] * inv_m[i]; }
    __device__ __forceinline__ double mc(int i, int r) const { return -S.M.mug - S.M.dd[r] * inv_m[i]; }

    // ---- per-problem setup --------------------------------------------------------------------
    __device__ void setup(const double* x0, const double* mass, const double* params, const int32_t* fm,
                          const double* Y = nullptr) {
        const int nl = S.nl, N = S.N, nv = S.nv, npv = S.npv;
        LANES(t, 2 * nl) pvec[t] = x0[t];
        LANES(t, S.npar) pvec[2 * nl + t] = params[t];
        if (lane == 0) pvec[npv - 1] = 1.0;
        LANES(i, nl) {
            inv_m[i] = rcp(mass[i]);
            v0[i] = x0[2 * i + 1];
            pc[i] = x0[2 * i] + x0[2 * i + 1];
            rlo[i * (N + 1)] = x0[2 * i + 1];
            rhi[i * (N + 1)] = x0[2 * i + 1];
        }
        __syncwarp(gm);
        double c = 0.0;
        bool infeas = false;
        if (Y) {
            // products with the shared structure matrices were done on the tensor cores (pm_precompute_kernel)
            LANES(j, nv) g0[j] = Y[j];
            LANES(r, S.ng) bgen[r] = Y[nv + r];
            LANES(r, S.n0) {
                const double s = Y[nv + S.ng + r];
                if (s > 0.0) {
                    if (isfinite(S.w0[r])) c += S.w0[r] * s;
                    else if (s > 1e-9) infeas = true;
                }
            }
            {
                const double* yr = Y + nv + S.ng + S.n0;
                LANES(r, S.nres) c += S.wres[r] * yr[r] * yr[r];
                LANES(l, S.nlin) c += yr[S.nres + l] * yr[S.nres + S.nlin + l];
            }
            yel = Y + S.o_yel;
            LANES(e, S.nel) c += 0.5 * yel[e] * yel[S.nel + e];     // - gE' C^-1 gE / 2 of the eliminated copies
        } else {
            // residual constants, param-linear factors, generic right-hand sides, constant rows
            LANES(r, S.nres) {
                double s = 0.0;
                _Pragma("unroll 4")
                for (int t = 0; t < npv; ++t) s += S.Cres[(size_t)r * npv + t] * pvec[t];
                cres[r] = s;
                c += S.wres[r] * s * s;
            }
            LANES(l, S.nlin) {
                double a = 0.0, p = 0.0;
                _Pragma("unroll 4")
                for (int t = 0; t < npv; ++t) {
                    a += S.La[(size_t)l * npv + t] * pvec[t];
                    p += S.Lp[(size_t)l * npv + t] * pvec[t];
                }
                cres[S.nres + l] = a;
                c += a * p;
            }
            LANES(r, S.ng) {
                double s = 0.0;
                _Pragma("unroll 4")
                for (int t = 0; t < npv; ++t) s += S.BR[(size_t)r * npv + t] * pvec[t];
                bgen[r] = s;
            }
            LANES(r, S.n0) {
                double s = 0.0;
                _Pragma("unroll 4")
                for (int t = 0; t < npv; ++t) s += S.B0[(size_t)r * npv + t] * pvec[t];
                if (s > 0.0) {
                    if (isfinite(S.w0[r])) c += S.w0[r] * s;
                    else if (s > 1e-9) infeas = true;
                }
            }
        }
        LANES(i, nl) {   // state row k = 1 on the (constant) position p_1 = p_0 + v_0
            if (pc[i] > S.pmax + 1e-9 || pc[i] < S.pmin - 1e-9) infeas = true;
        }
        __syncwarp(gm);
        c0 = wsum<GW>(gm, c);
        infeas = __any_sync(gm, infeas);
        if (!Y) {
            LANES(j, nv) {
                double s = 0.0;
                _Pragma("unroll 4")
                for (int r = 0; r < S.nres; ++r) s += S.RW2[(size_t)j * S.nres + r] * cres[r];
                _Pragma("unroll 4")
                for (int l = 0; l < S.nlin; ++l) s += S.Lz[(size_t)l * nv + j] * cres[S.nres + l];
                g0[j] = s;
            }
        }
        if (ONE) {                  // proximal centre: constant velocity
            LANES(j, nv) zc[j] = (j < nl * N) ? x0[2 * (j / N) + 1] : 0.0;
        }
        ppa = 0; retry = 0;
        built_L = -1;                      // H^-1 not loaded yet
        iters = nodes = it = q = 0;
        inc = HUGE_VAL; own = HUGE_VAL; trouble = limit = timeout = false; dive = true; sub_ord = 0;
        t_start = S.time_limit_ns > 0 ? hvp_now_ns() : 0;
        lev = 0; L = 0; fixed = FX;
        if (infeas) { state = PS_DONE; __syncwarp(gm); return; }
        if (FX) {
            bool bad = false;
            LANES(d, S.depth) {
                const int i = d % nl, k = d / nl, r = fm[i * N + k];
                modes[d] = r;
                if (r < 0 || r >= S.M.R) bad = true;
                else if (k == 0 && (v0[i] > S.M.hi[r] + 1e-9 || v0[i] < S.M.lo[r] - 1e-9)) bad = true;
            }
            bad = __any_sync(gm, bad);
            L = S.depth;
            state = bad ? PS_DONE : PS_BUILD;
        } else {
            if (lane == 0) { open_level(0); }
            state = PS_NEXT;
        }
        __syncwarp(gm);
    }

    // candidates of decision `lv` (lane 0 only)
    __device__ void open_level(int lv) {
        const int nl = S.nl, N = S.N;
        const int i = lv % nl, k = lv / nl;
        const double lo = rlo[i * (N + 1) + k], hi = rhi[i * (N + 1) + k];
        int cn = 0;
        _Pragma("unroll 1")
        for (int r = 0; r < S.M.R; ++r)
            if (S.M.lo[r] <= hi && S.M.hi[r] >= lo && S.M.lo[r] <= S.M.hi[r]) cn |= (1 << r);
        if (AD && lv < job_P) cn &= (1 << modes[lv]);        // adopted sub-tree: the donor's path, nothing beside it
        cand[lv] = cn;
        xstar[lv] = (k == 0) ? v0[i] : x[i * N + k - 1];
        pf[lv] = -HUGE_VAL;
    }

    // dual information of the node that has just been solved for the level that branches next (all lanes)
    __device__ void parent_info(double fnode) {
        const int nl = S.nl, N = S.N, nv = S.nv, ld = S.ld;
        const int i = lev % nl, k = lev / nl;
        if (k == 0 || ONE) return;                 // v0 is data; LP nodes: no curvature, bounds come from the LP
        const int j = i * N + k - 1;
        LANES(a, q) dv[a] = dot2(Nact + a * ld, 1, Hinv + j * ld, nv);          // d = N' H^-1 e_j
        __syncwarp(gm);
        double part = 0.0, tp = HUGE_VAL, tm = HUGE_VAL;
        LANES(a, q) {
            const double s = dot2(Ginv + a * ld, 1, dv, q);
            part += s * dv[a];
            const double la = lam[a], wm = soft_w(act[a]);
            if (s > 1e-14) {
                tp = fmin(tp, la * rcp(s));
                if (isfinite(wm)) tm = fmin(tm, (wm - la) * rcp(s));
            } else if (s < -1e-14) {
                tm = fmin(tm, la * rcp(-s));
                if (isfinite(wm)) tp = fmin(tp, (wm - la) * rcp(-s));
            }
        }
        const double nHn0 = Hinv[j * ld + j];
        double nz = nHn0 - wsum<GW>(gm, part);
        int dummy = 0;
        wargmin<GW>(gm, tp, dummy);
        wargmin<GW>(gm, tm, dummy);
        if (q == nv || !(nz > 1e-11 * nHn0)) nz = 0.0;      // dependent row: the dual is linear along the step
        if (lane == 0) { pf[lev] = fmin(fnode, dual); pnz[lev] = nz; ptp[lev] = fmax(tp, 0.0); ptm[lev] = fmax(tm, 0.0); }
        __syncwarp(gm);
    }

    // ---- NEXT: next node of the depth-first search (scalar; lane 0, result broadcast) ----------
    // Adoption, donor side (all lanes).  If workers are waiting, give away the shallowest untried sibling of the own
    // part of the depth-first stack (the largest piece of remaining work) to one of them: the candidate leaves cand[],
    // the waiting worker gets the path to it as a mode prefix and a result slot of its own.
    __device__ void try_donate() {
        const PmSplit& P = *sp;
        unsigned long long c = 0;
        if (lane == 0) c = *reinterpret_cast<volatile unsigned long long*>(P.ad_count);
        c = __shfl_sync(gm, c, 0, GW);
        if ((unsigned)c == 0u) return;
        int dl = -1;
        if (lane == 0) {
            const int top = min(lev, S.depth - 1 - P.ad_free);
            _Pragma("unroll 1")
            for (int l = (sub_M > 0 ? sub_D : 0); l <= top; ++l)
                if (cand[l]) { dl = l; break; }
            if (dl >= 0 && *reinterpret_cast<volatile int*>(P.pool_used) >= P.pool_cap) dl = -1;
        }
        dl = __shfl_sync(gm, dl, 0, GW);
        if (dl < 0) return;
        const int k = (ad_probe + lane) % ad_nwork;
        ad_probe = (ad_probe + GW + 1) % ad_nwork;
        const int s = *reinterpret_cast<volatile int*>(P.mail_state + k);
        const unsigned hit = __ballot_sync(gm, s == 1) & gm;
        if (!hit) return;
        const int kk = __shfl_sync(gm, k, (__ffs(hit) - 1) & (GW - 1), GW);
        if (lane == 0 && atomicCAS(P.mail_state + kk, 1, 3) == 1) {
            const int slot = atomicAdd(P.pool_used, 1);
            if (slot >= P.pool_cap) {
                atomicExch(P.mail_state + kk, 1);
            } else {
                atomicAdd(P.ad_count, ~0ull);                  // one waiting worker less
                const int rg = __ffs(cand[dl]) - 1;
                cand[dl] &= ~(1 << rg);
                int* job = P.mail_job + (size_t)kk * P.mail_stride;
                job[0] = fidx; job[1] = dl + 1; job[2] = slot;
                _Pragma("unroll 1")
                for (int l = 0; l < dl; ++l) job[3 + l] = modes[l];
                job[3 + dl] = rg;
                P.pool_owner[slot] = fidx;
                __threadfence();
                atomicExch(P.mail_state + kk, 2);
            }
        }
        __syncwarp(gm);
    }

    __device__ void do_next() {
        if (AD && ad_nwork > 0 && !dive) try_donate();
        int st = PS_DONE, nlev = lev, nL = L;
        if (lane == 0) {
            const double eps = 1e-9;
            const int nl = S.nl, N = S.N;
            _Pragma("unroll 1")
            for (;;) {
                const int cset = cand[nlev];
                if (cset == 0) {
                    if (nlev == 0) { st = PS_DONE; break; }
                    --nlev;
                    continue;
                }
                const int i = nlev % nl, k = nlev / nl;
                int rg = -1; double bd = HUGE_VAL;
                const double xs = xstar[nlev];
                if (sub_M > 0 && nlev < sub_D) {
                    rg = __ffs(cset) - 1;      // prefix levels: a fixed order, identical in every warp of the problem
                } else {
                    _Pragma("unroll 1")
                    for (int c = 0; c < S.M.R; ++c) {
                        if (!((cset >> c) & 1)) continue;
                        const double lo = S.M.lo[c], hi = S.M.hi[c];
                        const double dist = xs < lo ? lo - xs : (xs > hi ? xs - hi : 0.0);
                        if (dist < bd) { bd = dist; rg = c; }
                    }
                }
                cand[nlev] = cset & ~(1 << rg);
                if (bd > 0.0 && inc < HUGE_VAL && pf[nlev] > -HUGE_VAL) {
                    // sibling bound from the solved parent (parent_info)
                    const double nz = pnz[nlev];
                    const double tmax = (xs > S.M.hi[rg]) ? ptp[nlev] : ptm[nlev];
                    const double t = (nz > 0.0) ? fmin(bd * rcp(nz), tmax) : tmax;
                    const double bound = (t < HUGE_VAL) ? pf[nlev] + t * bd - 0.5 * t * t * nz : HUGE_VAL;
                    if (bound > inc) continue;
                }
                modes[nlev] = rg;
                const double jlo = fmax(rlo[i * (N + 1) + k], S.M.lo[rg]), jhi = fmin(rhi[i * (N + 1) + k], S.M.hi[rg]);
                if (jlo > jhi + eps) continue;
                const double a = ma(i, rg), b = mb(i, rg), c = mc(i, rg);
                double nlo = fmax(a * jlo + c + b * S.umin, jlo + S.a_dec + k * S.tight);
                double nhi = fmin(a * jhi + c + b * S.umax, jhi + S.a_acc - k * S.tight);
                nlo = fmax(nlo, S.vmin); nhi = fmin(nhi, S.vmax);
                if (nlo > nhi + eps) continue;
                rlo[i * (N + 1) + k + 1] = nlo - eps; rhi[i * (N + 1) + k + 1] = nhi + eps;
                nL = nlev + 1;
                if (sub_M > 0) {
                    if (nL < sub_D) {          // mode prefix not complete yet: descend without solving
                        ++nlev;
                        open_level(nlev);
                        continue;
                    }
                    if (nL == sub_D && (sub_ord++ % sub_M) != sub_code) continue;   // another warp's sub-tree
                }
                if (dive && nodes >= 1 && nL < S.depth) {
                    // first descent: without an incumbent the relaxations along the path cannot prune, they
                    // only guide the region choice -- follow the last relaxed trajectory and solve the leaf
                    ++nlev;
                    open_level(nlev);
                    continue;
                }
                st = PS_BUILD;
                break;
            }
        }
        __syncwarp(gm);
        state = __shfl_sync(gm, st, 0, GW);
        lev = __shfl_sync(gm, nlev, 0, GW);
        L = __shfl_sync(gm, nL, 0, GW);
        if (state == PS_BUILD) hull();
    }

    // Interval reachability of the NOT yet fixed stages (lane v = vehicle v): from the velocity interval
    // at a vehicle's first unfixed stage, the hull over all PWA modes it can be in gives, stage by stage,
    // a velocity interval and the acceleration range its traction / braking allows.  Every completion of
    // the node's mode prefix satisfies them, so the node relaxation may impose them (they replace the
    // constant a_dec / a_acc rows of the unfixed stages) and stays a valid lower bound.
    __device__ void hull() {
        const int nl = S.nl, N = S.N;
        const double eps = 1e-9;
        LANES(v, nl) {
            int kf = (L - v + nl - 1) / nl;              // stages 0..kf-1 of vehicle v are fixed
            if (kf < 0) kf = 0;
            double lo = rlo[v * (N + 1) + kf], hi = rhi[v * (N + 1) + kf];
            _Pragma("unroll 1")
            for (int s = kf; s < N; ++s) {
                double nlo = HUGE_VAL, nhi = -HUGE_VAL, dmax = -HUGE_VAL, dmin = HUGE_VAL;
                _Pragma("unroll 1")
                for (int r = 0; r < S.M.R; ++r) {
                    const double jlo = fmax(lo, S.M.lo[r]), jhi = fmin(hi, S.M.hi[r]);
                    if (jlo > jhi + eps) continue;
                    const double a = ma(v, r), b = mb(v, r), c = mc(v, r);
                    nlo = fmin(nlo, a * jlo + c + b * S.umin);
                    nhi = fmax(nhi, a * jhi + c + b * S.umax);
                    dmax = fmax(dmax, (a - 1.0) * jlo + c + b * S.umax);     // a < 1: largest at the low end
                    dmin = fmin(dmin, (a - 1.0) * jhi + c + b * S.umin);
                }
                const double acc = S.a_acc - s * S.tight, dec = S.a_dec + s * S.tight;
                dmax = fmin(dmax + eps, acc); dmin = fmax(dmin - eps, dec);
                nlo = fmax(fmax(nlo, lo + dec), S.vmin); nhi = fmin(fmin(nhi, hi + acc), S.vmax);
                amax[v * N + s] = dmax; amin[v * N + s] = dmin;
                lo = nlo - eps; hi = nhi + eps;
                rlo[v * (N + 1) + s + 1] = lo; rhi[v * (N + 1) + s + 1] = hi;
            }
        }
        __syncwarp(gm);
    }

    // st: 0 solved, 1 infeasible, 2 numerical trouble
    __device__ void node_done(int st, double obj) {
        iters += it;
        if (ONE) {
            // An LP node is degenerate far more often than a QP node (a full active set is the rule, ties between a
            // partial and a full step are common): a round that ends in numerical trouble is repeated from a slightly
            // different proximal centre, which changes the path of the active-set method but not the LP it converges to.
            if (st == 2 && retry < 4) {
                ++retry;
                LANES(j, S.nv) {
                    const unsigned h = (unsigned)(j * 2654435761u) ^ (unsigned)(retry * 40503u) ^ (unsigned)(nodes * 9973u);
                    zc[j] += 1e-3 * ((double)((h >> 8) & 0xffff) / 65536.0 - 0.5) * (double)retry;
                }
                ppa = 0;
                state = PS_BUILD;
                __syncwarp(gm);
                return;
            }
            retry = 0;
        }
        ++nodes;
        state = FX ? PS_DONE : PS_NEXT;
        if (st != 0 || L == S.depth) dive = false;
        // the node budget of a probing wave counts EVERY node: checked only on the way down (as it was), a probe with a
        // budget of 8 went on for 30-70 nodes of leaves and pruned nodes -- 24 / 52 ms at n = 8 / 10, as long as the search
        if (stop_nodes > 0 && nodes >= stop_nodes && !FX) { limit = true; state = PS_DONE; }
        if (st == 2) { trouble = true; return; }
        if (st == 1) return;
        if (inc < HUGE_VAL && !(obj < hvp_cut(inc, S.mip_gap))) return;              // bound
        if (L == S.depth) {                                                           // leaf
            inc = obj; own = obj;
            LANES(j, S.nv) best[j] = x[j];
            LANES(d, S.depth) bmodes[d] = modes[d];
            if (shared && lane == 0) atomicMin(shared, dkey(obj));
            __syncwarp(gm);
            return;
        }
        if (state == PS_DONE) return;
        if (S.max_nodes > 0 && nodes >= S.max_nodes) { limit = true; state = PS_DONE; return; }
        if (S.time_limit_ns > 0) {          // warp-uniform decision: lane 0 reads the clock
            long long now = 0;
            if (lane == 0) now = hvp_now_ns();
            now = __shfl_sync(gm, now, 0, GW);
            if (now - t_start > S.time_limit_ns) { timeout = true; state = PS_DONE; return; }
        }
        if (stop_nodes > 0 && nodes >= stop_nodes) { limit = true; state = PS_DONE; return; }
        if (budget > 0 && nodes >= budget) {
            // heavy tree: hand it to the sub-tree pass if the list has room, else finish it here
            int slot = 0;
            if (lane == 0) slot = atomicAdd(sp->nflag, 1);
            slot = __shfl_sync(gm, slot, 0, GW);
            if (slot < sp->cap) {
                if (lane == 0) { sp->flagged[slot] = (int)prob; sp->inc_shared[slot] = dkey(inc); }
                limit = true; state = PS_DONE;
                return;
            }
            budget = 0;
        }
        ++lev;
        if (lane == 0) open_level(lev);
        __syncwarp(gm);
        static_assert(true, "");
        if (S.sibling) parent_info(obj);
    }

    // L1 weight of a row id (+inf: hard row).  Soft rows: generic rows with a finite wmax, and the |u| pair of a fixed stage.
    __device__ __forceinline__ double soft_w(int id) const {
        const int pt = id / 4096, ix = id % 4096;
        if (pt == PT_GEN) return S.wmax[ix];
        if (pt == PT_USP || pt == PT_USN) return S.qu * rcp(bm[(ix % S.N) * S.nl + ix / S.N]);
        return HUGE_VAL;
    }
    // a saturated soft row changes orientation (its penalty gradient stays in x, which is only ever updated incrementally)
    __device__ __forceinline__ void flip(int id) {
        const int pt = id / 4096, ix = id % 4096;
        if (pt == PT_GEN) orient[ix] = -orient[ix];
        else if (pt == PT_USP) uor[ix] = -uor[ix];
        else if (pt == PT_USN) uor[S.nv + ix] = -uor[S.nv + ix];
    }

    // H <- H + s * 2 qu e e',  e = (e_jk - a e_jm)/b  (decision d in mode rg): Sherman-Morrison on H^-1
    __device__ void rank1(int d, int rg, double s) {
        const int nl = S.nl, N = S.N, nv = S.nv, ld = S.ld;
        const int i = d % nl, k = d / nl, jk = i * N + k, jm = jk - 1;
        const double ib = rcp(mb(i, rg)), ea = (k >= 1) ? -ma(i, rg) * ib : 0.0;
        LANES(j, nv) {
            double v = ib * Hinv[j * ld + jk];
            if (k >= 1) v += ea * Hinv[j * ld + jm];
            wv[j] = v;
        }
        __syncwarp(gm);
        const double ev = ib * wv[jk] + ((k >= 1) ? ea * wv[jm] : 0.0);
        const double den = rcp(s * (0.5 / S.qu) + ev);
        LANES(j, nv) {
            const double vj = wv[j] * den;
            _Pragma("unroll 4")
            for (int c = 0; c < nv; ++c) Hinv[j * ld + c] -= vj * wv[c];
        }
        __syncwarp(gm);
    }

    // ---- BUILD: bring H^-1 to this node, gradient, unconstrained minimiser ----------------------
    __device__ void do_build() {
        const int nl = S.nl, N = S.N, nv = S.nv, ld = S.ld;
        if (shared) {       // best objective found by ANY warp working on this problem
            double g = 0.0;
            if (lane == 0) g = dval(*reinterpret_cast<volatile unsigned long long*>(shared));
            g = __shfl_sync(gm, g, 0, GW);
            if (g < inc) inc = g;
        }
        int c = 0;
        if (built_L > 0)
            while (c < built_L && c < L && built[c] == modes[c]) ++c;
        const bool reload = built_L < 0 || c == 0 || (built_L - c > c);
        __syncwarp(gm);
        if (reload) {
            _Pragma("unroll 1")
            for (int e = lane; e < nv * nv; e += GW) Hinv[(e / nv) * ld + (e % nv)] = S.H0inv[e];
            c = 0; built_L = 0;
            __syncwarp(gm);
        }
        if (!ONE) {                 // 1-norm: no quadratic input cost, H^-1 = I / rho_px for every node
            _Pragma("unroll 1")
            for (int d = built_L - 1; d >= c; --d) rank1(d, built[d], -1.0);
            _Pragma("unroll 1")
            for (int d = c; d < L; ++d) rank1(d, modes[d], 1.0);
        }
        LANES(d, L) {
            built[d] = modes[d];
            const int i = d % nl, r = modes[d];
            am[d] = ma(i, r); bm[d] = mb(i, r); cm[d] = mc(i, r);
        }
        built_L = L;
        __syncwarp(gm);
        const double qu = S.qu;
        // gradient of the node's smooth part: 2-norm -- g0 + input-cost terms of the fixed stages; 1-norm -- the proximal
        // term plus the penalty gradients w n of the soft rows that are currently in their saturated orientation
        auto gradient = [&]() {
            LANES(j, nv) {
                double g = g0[j];
                if (ONE) {
                    g -= S.rho_px * zc[j];          // gradient of rho/2 |z - zc|^2 at 0
                    _Pragma("unroll 1")
                    for (int r = 0; r < S.ng; ++r)
                        if (orient[r] < 0) g += S.wmax[r] * S.AT[(size_t)j * S.ng + r];
                    if (j < nl * N) {
                        const int i = j / N, kk = j % N;
                        const int d0 = kk * nl + i, d1 = d0 + nl;
                        if (d0 < L) g += qu * rcp(bm[d0]) * (double)(uor[nv + j] - uor[j]) * 0.5;              // +-1 on own stage
                        if (kk + 1 < N && d1 < L) g -= qu * rcp(bm[d1]) * am[d1] * (double)(uor[nv + j + 1] - uor[j + 1]) * 0.5;
                    }
                } else if (j < nl * N) {
                    const int i = j / N, kk = j % N;
                    const int d0 = kk * nl + i, d1 = d0 + nl;
                    if (d0 < L) {
                        const double ib = rcp(bm[d0]);
                        const double kc = (kk == 0) ? -(am[d0] * v0[i] + cm[d0]) * ib : -cm[d0] * ib;
                        g += 2.0 * qu * kc * ib;
                    }
                    if (kk + 1 < N && d1 < L) {
                        const double ib = rcp(bm[d1]);
                        g += 2.0 * qu * (-cm[d1] * ib) * (-am[d1] * ib);
                    }
                }
                gn[j] = g;
            }
            __syncwarp(gm);
        };
        // WARM proximal round (1-norm, rounds after the first of a node): the centre has moved to the last solution x,
        // everything else -- active rows, their normals, (N'H^-1N)^-1, orientations -- is still in place.  New
        // unconstrained minimiser x_u, residuals of the active rows c = N'(x_u - x) (they were zero at x), multipliers
        // lambda = Ginv c; if all of them are admissible this is the solution on the same active set and SELECT goes on
        // from there -- at the fixed point that is the whole round.  Otherwise: the cold start below.
        if (ONE && ppa > 0 && q > 0) {
            gradient();
            double xs_ = 0.0; int xi_ = 0;
            LANES(j, nv) { const double v_ = -dot2(Hinv + j * ld, 1, gn, nv); wv[j] = v_; xs_ = fmax(xs_, fabs(v_)); }          // x_u
            wargmax<GW>(gm, xs_, xi_);
            xu_scale = xs_;
            __syncwarp(gm);
            LANES(a, q) dv[a] = dot2(Nact + a * ld, 1, wv, nv) - dot2(Nact + a * ld, 1, x, nv);
            __syncwarp(gm);
            bool ok = true;
            LANES(a, q) {
                const double l = dot2(Ginv + a * ld, 1, dv, q);
                rv[a] = l;
                if (!(l >= 0.0) || l > soft_w(act[a])) ok = false;
            }
            ok = __all_sync(gm, ok);
            if (ok) {
                LANES(a, q) lam[a] = rv[a];
                __syncwarp(gm);
                LANES(j, nv) np_[j] = dot2(Nact + j, ld, lam, q);          // N lambda
                __syncwarp(gm);
                LANES(j, nv) x[j] = wv[j] - dot2(Hinv + j * ld, 1, np_, nv);
                it = 0;
                state = PS_SELECT;
                __syncwarp(gm);
                return;
            }
            __syncwarp(gm);
        }
        // orientation of the soft rows.  2-norm: every L1 row starts in its plain orientation.  1-norm: a cost term whose
        // row is violated AT THE PROXIMAL CENTRE starts saturated (orientation -1, its penalty gradient w n in the
        // objective) -- near the solution that is where it ends up, so a round needs about one active-set step per KINK
        // instead of one per cost term.
        LANES(r, S.ng) {
            int o = 1;
            if (ONE && isfinite(S.wmax[r]) && dot2(S.AT + r, S.ng, zc, nv) - bgen[r] > 0.0) o = -1;
            orient[r] = o; agen[r] = 0;
        }
        LANES(j, nv) {
            aflag[j] = 0;
            int op = 1, on = 1;
            if (ONE && j < nl * N) {
                const int i = j / N, kk = j % N, d0 = kk * nl + i;
                if (d0 < L) {
                    const double du = zc[j] - am[d0] * (kk == 0 ? v0[i] : zc[j - 1]) - cm[d0];
                    if (du > 0.0) op = -1; else if (du < 0.0) on = -1;
                }
            }
            uor[j] = op; uor[nv + j] = on;
        }
        __syncwarp(gm);
        gradient();
        double dpart = 0.0, xs_ = 0.0;
        LANES(j, nv) {
            const double s = dot2(Hinv + j * ld, 1, gn, nv);
            x[j] = -s;
            dpart -= 0.5 * gn[j] * s;
            if (ONE) xs_ = fmax(xs_, fabs(s));
        }
        if (ONE) { int xi_ = 0; wargmax<GW>(gm, xs_, xi_); xu_scale = xs_; }
        LANES(d, L) {
            const int i = d % nl, k = d / nl;
            const double kc = ((k == 0) ? -(am[d] * v0[i] + cm[d]) : -cm[d]) * rcp(bm[d]);
            dpart += qu * kc * kc;
        }
        // value of the node's dual function at the unconstrained minimiser; it only grows from here
        dual = c0 + wsum<GW>(gm, dpart);
        it = 0; q = 0;
        state = PS_SELECT;
        __syncwarp(gm);
    }

    // most violated row of the node QP at x (rows already active have residual ~0); `soft` = also
    // consider the L1-penalised rows (in their current orientation)
    __device__ void scan(double& best_v, int& bid, double tol, bool soft) {
        const int nl = S.nl, N = S.N, nv = S.nv, ng = S.ng;
        best_v = tol; bid = 0x7fffffff;
        // rows that are in the active set are skipped: their residual is zero up to round-off, which in
        // heavily constrained, ill-conditioned nodes can exceed `tol` and must not be re-selected
#define PM_CAND(T, IDX, SV)                                                        \
    {                                                                              \
        const double s__ = (SV);                                                   \
        if (s__ > best_v && !((af__ >> (T)) & 1)) { best_v = s__; bid = PM_ID(T, IDX); } \
    }
        LANES(j, nl * N) {
            const int af__ = aflag[j];
            const int i = j / N, kk = j % N;
            const int d0 = kk * nl + i, d1 = d0 + nl;
            const double xv = x[j];
            double lo = S.vmin, hi = S.vmax;
            if (kk + 1 < N && d1 < L) {
                const int r = modes[d1];
                lo = fmax(lo, S.M.lo[r]); hi = fmin(hi, S.M.hi[r]);
            }
            if (d0 >= L) {       // stage kk not fixed: reachable interval of v_{kk+1}
                lo = fmax(lo, rlo[i * (N + 1) + kk + 1]); hi = fmin(hi, rhi[i * (N + 1) + kk + 1]);
            }
            if (ONE && soft && d0 < L) {      // qu |u| of a fixed stage: the pair of soft rows on b u
                const double du = xv - am[d0] * (kk == 0 ? v0[i] : x[j - 1]) - cm[d0];
                PM_CAND(PT_USP, j, uor[j] > 0 ? du : -du);
                PM_CAND(PT_USN, j, uor[nv + j] > 0 ? -du : du);
            }
            if (kk == 0) {
                lo = fmax(lo, v0[i] + S.a_dec); hi = fmin(hi, v0[i] + S.a_acc);
                if (d0 < L) {
                    const double m = am[d0] * v0[i] + cm[d0];
                    lo = fmax(lo, m + bm[d0] * S.umin); hi = fmin(hi, m + bm[d0] * S.umax);
                }
            } else {
                const double xm = x[j - 1];
                double acc = S.a_acc - kk * S.tight, dec = S.a_dec + kk * S.tight;
                if (d0 < L) {
                    const double du = xv - am[d0] * xm - cm[d0];
                    PM_CAND(PT_UHI, j, du - bm[d0] * S.umax);
                    PM_CAND(PT_ULO, j, bm[d0] * S.umin - du);
                } else {
                    acc = amax[j]; dec = amin[j];
                }
                const double dvv = xv - xm;
                PM_CAND(PT_ACC, j, dvv - acc);
                PM_CAND(PT_DEC, j, dec - dvv);
            }
            PM_CAND(PT_UB, j, xv - hi);
            PM_CAND(PT_LB, j, lo - xv);
        }
        if (N >= 2) {
            LANES(i, nl) {
                const int af__ = aflag[i];
                double ps = pc[i];
                PM_CAND(PT_PLO, i, S.pmin - (ps + x[i * N]));
                _Pragma("unroll 1")
                for (int kk = 0; kk + 1 < N; ++kk) ps += x[i * N + kk];
                PM_CAND(PT_PHI, i, ps - S.pmax);
            }
        }
        LANES(r, ng) {
            if (!soft && isfinite(S.wmax[r])) continue;
            const int af__ = agen[r] << PT_GEN;
            const double s = dot2(S.AT + r, ng, x, nv) - bgen[r];
            PM_CAND(PT_GEN, r, orient[r] > 0 ? s : -s);
        }
#undef PM_CAND
        wargmax<GW>(gm, best_v, bid);
    }

    // objective of the node QP at x (tracking + input cost of the fixed stages + L1 penalties)
    __device__ double objective() {
        const int nl = S.nl, N = S.N, nv = S.nv, ng = S.ng;
        double f = 0.0;
        if (!ONE) {
            LANES(j, nv) {
                const double s = dot2(S.H0 + (size_t)j * nv, 1, x, nv);
                f += x[j] * (g0[j] + 0.5 * s);
            }
        }
        LANES(d, L) {
            const int i = d % nl, k = d / nl, jk = i * N + k;
            const double xp = (k >= 1) ? x[jk - 1] : v0[i];
            const double uu = (x[jk] - am[d] * xp - cm[d]) * rcp(bm[d]);
            f += ONE ? S.qu * fabs(uu) : S.qu * uu * uu;     // the proximal term is not part of the LP objective
        }
        LANES(r, ng) {
            const double wm = S.wmax[r];
            if (isfinite(wm)) {
                const double s = dot2(S.AT + r, ng, x, nv) - bgen[r];
                if (s > 0.0) f += wm * s;
            }
        }
        return c0 + wsum<GW>(gm, f);
    }

    // 1-norm: one step of iterative refinement on the active rows.  With H = rho I (rho = 1e-3) and multipliers up to
    // w = 1e4 the point x = -H^-1 (g + N lambda) carries a round-off of ~1e-8, which a kink of weight w turns into 1e-4
    // of objective.  The minimum-norm correction dx = -H^-1 N (N'H^-1 N)^-1 (N'x - b) puts x back on its active rows.
    __device__ void refine() {
        const int nv = S.nv, ld = S.ld;
        if (q == 0) return;
        LANES(a, q) dv[a] = dot2(Nact + a * ld, 1, x, nv) - zd[a];
        __syncwarp(gm);
        LANES(a, q) rv[a] = dot2(Ginv + a * ld, 1, dv, q);
        __syncwarp(gm);
        LANES(j, nv) np_[j] = dot2(Nact + j, ld, rv, q);
        __syncwarp(gm);
        LANES(j, nv) x[j] -= dot2(Hinv + j * ld, 1, np_, nv);
        __syncwarp(gm);
    }

    // ---- SELECT: most violated row, or the node is solved --------------------------------------
    __device__ void do_select() {
        const int nl = S.nl, N = S.N, nv = S.nv, ng = S.ng;
        double best_v; int bid;
        scan(best_v, bid, 1e-9, true);
        if (bid == 0x7fffffff) {
            if (ONE) {
                // Proximal-point round done.  The node LP is solved once the centre stops moving -- or, when the
                // optimum is a face and round-off keeps the point wandering on it, once the LP objective stops
                // decreasing: f(zc) - f(x) >= rho/2 |x - zc|^2, so a stalled objective means a stalled point, and a
                // proximal step from a non-optimal centre always decreases f.
                double dl = 0.0; int dj = 0;
                LANES(j, nv) { const double a = fabs(x[j] - zc[j]); if (a > dl) dl = a; }
                wargmax<GW>(gm, dl, dj);
                const double f = objective();
                // x = x_u - H^-1 N lambda is a difference of vectors of size |x_u|: with penalty gradients of 1e5 (a platoon
                // deep inside its safety distance) and rho = 1e-3 that is 1e8, and the iterates jitter by ~1e-6 for ever --
                // 5 of 49 152 per-vehicle MILPs ended as "numerical trouble" at the correct optimum.  The movement test
                // therefore scales with the round-off level of the round; refine() then puts x on its active rows.
                const double dtol = fmax(1e-6, 1e-13 * xu_scale);
                const bool stalled = ppa >= 1 && f_prev - f <= S.ppa_stall * fmax(1.0, fabs(f));
                if (dl > dtol && !stalled && ppa < 200) {
                    LANES(j, nv) zc[j] = x[j];
                    ++ppa;
                    f_prev = f;
                    iters += it;
                    state = PS_BUILD;
                    __syncwarp(gm);
                    return;
                }
                if (dl > dtol && !stalled) { node_done(2, 0.0); return; }
                ppa = 0;
                refine();
                node_done(0, objective());
                return;
            }
            ppa = 0;
            node_done(0, objective());
            return;
        }
        pid = bid;
        const int pt = pid / 4096, idx = pid % 4096;
        p_soft = isfinite(soft_w(pid));
        // dense normal of p
        LANES(j, nv) {
            double v = 0.0;
            switch (pt) {
                case PT_UB: v = (j == idx) ? 1.0 : 0.0; break;
                case PT_LB: v = (j == idx) ? -1.0 : 0.0; break;
                case PT_UHI: case PT_ULO: {
                    const int i = idx / N, kk = idx % N, d0 = kk * nl + i;
                    v = (j == idx) ? 1.0 : ((j == idx - 1) ? -am[d0] : 0.0);
                    if (pt == PT_ULO) v = -v;
                } break;
                case PT_USP: case PT_USN: {
                    const int i = idx / N, kk = idx % N, d0 = kk * nl + i;
                    v = (j == idx) ? 1.0 : ((j == idx - 1 && kk > 0) ? -am[d0] : 0.0);
                    v *= (pt == PT_USP) ? (double)uor[idx] : -(double)uor[nv + idx];
                } break;
                case PT_ACC: v = (j == idx) ? 1.0 : ((j == idx - 1) ? -1.0 : 0.0); break;
                case PT_DEC: v = (j == idx) ? -1.0 : ((j == idx - 1) ? 1.0 : 0.0); break;
                case PT_PHI: v = (j >= idx * N && j < idx * N + N - 1) ? 1.0 : 0.0; break;
                case PT_PLO: v = (j == idx * N) ? -1.0 : 0.0; break;
                default: v = S.AT[(size_t)j * ng + idx] * (double)orient[idx]; break;
            }
            np_[j] = v;
        }
        __syncwarp(gm);
        // yp = H^-1 n_p, nHn = n_p' yp
        double acc = 0.0;
        LANES(j, nv) {
            const double s = dot2(Hinv + j * S.ld, 1, np_, nv);
            yp[j] = s;
            acc += s * np_[j];
        }
        nHn = wsum<GW>(gm, acc);
        if (ONE) {          // right-hand side of p in its current orientation (best_v = n_p'x - b_p), kept for refine()
            double nx = 0.0;
            LANES(j, nv) nx += np_[j] * x[j];
            bp = wsum<GW>(gm, nx) - best_v;
        }
        cp = best_v;
        lam_p = 0.0;
        state = PS_STEP;
        __syncwarp(gm);
    }

    // ---- STEP: one primal-dual step towards adding p -------------------------------------------
    __device__ void do_step() {
        const int nv = S.nv, ld = S.ld;
        const double tol = 1e-9, INF = HUGE_VAL;
        if (++it > (ONE ? 200 * nv + 2000 : 40 * nv + 200)) { node_done(2, 0.0); return; }
        // p can reach its boundary exactly at the end of a PARTIAL step (t1 = t2 tie): it then joins the
        // active set with the multiplier it has accumulated (a full step of length zero) -- returning to
        // SELECT here would drop lam_p * n_p from the stationarity condition
        const bool zero_step = (cp <= tol);
        if (zero_step && !(lam_p > 0.0)) { state = PS_SELECT; return; }
        // d = N' yp
        LANES(a, q) {
            dv[a] = dot2(Nact + a * ld, 1, yp, nv);
        }
        __syncwarp(gm);
        // r = Ginv d ; nz = nHn - d'r ; ratio tests
        double part = 0.0, t1 = INF, t3 = INF;
        int k1 = 0x7fffffff, k3 = 0x7fffffff;
        LANES(a, q) {
            const double s = dot2(Ginv + a * ld, 1, dv, q);
            rv[a] = s;
            part += s * dv[a];
            const double la = lam[a];
            if (s > 1e-14) {
                const double t = la * rcp(s);
                if (t < t1) { t1 = t; k1 = a; }
            } else if (s < -1e-14) {
                const double wm = soft_w(act[a]);
                if (isfinite(wm)) {
                    const double t = (wm - la) * rcp(-s);
                    if (t < t3) { t3 = t; k3 = a; }
                }
            }
        }
        const double nz = nHn - wsum<GW>(gm, part);
        wargmin<GW>(gm, t1, k1);
        wargmin<GW>(gm, t3, k3);
        const bool dependent = (q == nv) || !(nz > 1e-11 * nHn);
        if (zero_step && dependent) { node_done(2, 0.0); return; }
        const double t2 = dependent ? INF : (zero_step ? 0.0 : cp * rcp(nz));
        const double t3p = p_soft ? (soft_w(pid) - lam_p) : INF;
        const double t = fmin(fmin(t1, t2), fmin(t3, t3p));
        if (!(t < INF)) { node_done(1, 0.0); return; }          // infeasible node
        // d(dual)/dt = violation of p along the step: the dual value is a lower bound on the node optimum
        dual += t * cp - (dependent ? 0.0 : 0.5 * t * t * nz);
        // (not under the 1-norm cost: the dual of a proximal sub-problem does not bound the node LP)
        if (!ONE && dual > inc) { node_done(1, 0.0); return; }          // the node cannot beat the incumbent
        __syncwarp(gm);
        if (!dependent) {
            // w = n_p - N r ;  x -= t H^-1 w ;  the violation of p shrinks by t nz
            LANES(j, nv) {
                wv[j] = np_[j] - dot2(Nact + j, ld, rv, q);
            }
            __syncwarp(gm);
            LANES(j, nv) {
                x[j] -= t * dot2(Hinv + j * ld, 1, wv, nv);
            }
            cp -= t * nz;
        }
        LANES(a, q) lam[a] -= t * rv[a];
        lam_p += t;
        if (t == t2) {
            // p becomes slot q: bordering update of Ginv with Schur complement nz
            const double is = rcp(nz);
            LANES(a, q) {
                const double ra = rv[a] * is;
                _Pragma("unroll 4")
                for (int b = 0; b < q; ++b) Ginv[a * ld + b] += ra * rv[b];
                Ginv[a * ld + q] = -ra;
                Ginv[q * ld + a] = -ra;
            }
            LANES(j, nv) Nact[q * ld + j] = np_[j];
            if (lane == 0) {
                Ginv[q * ld + q] = is;
                act[q] = pid;
                lam[q] = lam_p;
                if (ONE) zd[q] = bp;
                const int pt_ = pid / 4096, ix_ = pid % 4096;
                if (pt_ == PT_GEN) agen[ix_] = 1; else aflag[ix_] |= (1 << pt_);
            }
            ++q;
            state = PS_SELECT;
            __syncwarp(gm);
            return;
        }
        if (t == t3p) {                                          // soft p saturates: flip, not added
            if (lane == 0) flip(pid);
            state = PS_SELECT;
            __syncwarp(gm);
            return;
        }
        int drop;
        if (t == t1) drop = k1;
        else {                                                   // active soft row saturates: flip + drop
            drop = k3;
            if (lane == 0) flip(act[drop]);
        }
        if (lane == 0) {
            const int id_ = act[drop], pt_ = id_ / 4096, ix_ = id_ % 4096;
            if (pt_ == PT_GEN) agen[ix_] = 0; else aflag[ix_] &= ~(1 << pt_);
        }
        __syncwarp(gm);
        {   // Ginv <- Ginv - g g'/g_dd on the remaining slots, then move the last slot into `drop`
            const double idd = rcp(Ginv[drop * ld + drop]);
            LANES(a, q) dv[a] = Ginv[a * ld + drop];
            __syncwarp(gm);
            LANES(a, q) {
                const double da = dv[a] * idd;
                _Pragma("unroll 4")
                for (int b = 0; b < q; ++b) Ginv[a * ld + b] -= da * dv[b];
            }
            __syncwarp(gm);
            const int last = q - 1;
            if (drop != last) {
                LANES(b, q) Ginv[drop * ld + b] = Ginv[last * ld + b];
                __syncwarp(gm);
                LANES(a, q) Ginv[a * ld + drop] = Ginv[a * ld + last];
                __syncwarp(gm);
                if (lane == 0) {
                    Ginv[drop * ld + drop] = Ginv[last * ld + last];
                    act[drop] = act[last];
                    lam[drop] = lam[last];
                    if (ONE) zd[drop] = zd[last];
                }
                LANES(j, nv) Nact[drop * ld + j] = Nact[last * ld + j];
            }
            --q;
            __syncwarp(gm);
        }
        // state stays PS_STEP: continue with the same p
    }

    __device__ void solve() {
        while (state != PS_DONE) {
            if (!FX && state == PS_NEXT) do_next();
            if (state == PS_BUILD) do_build();
            if (state == PS_SELECT) do_select();
            if (state == PS_STEP) do_step();
        }
    }

    // ---- eval_cost: objective of a pinned (x[:, :N], u) guess, +inf if it is not feasible for the
    // MLD model within Gurobi's feasibility tolerance (fleet_event_based.py:308-327).  The modes are
    // whatever makes the pinned data consistent; the free last state follows from the dynamics, with
    // the cheaper mode when the last pinned velocity sits on a region boundary.
    __device__ double eval(const double* xg, const double* ug) {
        const int nl = S.nl, N = S.N, np1 = N + 1;
        const double tol = 1e-6;
        if (state == PS_DONE) return HUGE_VAL;                    // constant rows already infeasible
        bool bad = false;
        int lastm[4] = {-1, -1, -1, -1};                             // candidate modes of the last stage (lane = vehicle)
        int ncm = 0;
        LANES(i, nl) {
            const double* xp = xg + (size_t)i * 2 * np1;
            const double* xv = xp + np1;
            _Pragma("unroll 1")
            for (int k = 0; k < N; ++k) {
                const double v = xv[k], uu = ug[i * N + k];
                if (k + 1 < N && fabs(xp[k + 1] - (xp[k] + v)) > tol) bad = true;
                int found = -1;
                _Pragma("unroll 1")
                for (int r = 0; r < S.M.R; ++r) {
                    if (!(v >= S.M.lo[r] - tol && v <= S.M.hi[r] + tol)) continue;
                    if (k + 1 < N && fabs(xv[k + 1] - (ma(i, r) * v + mc(i, r) + mb(i, r) * uu)) > tol) continue;
                    if (found < 0) found = r;
                    if (k + 1 == N && ncm < 4) {                  // free last state: every admissible mode counts
                        if (ncm == 0) lastm[0] = r; else if (ncm == 1) lastm[1] = r; else if (ncm == 2) lastm[2] = r; else lastm[3] = r;
                        ++ncm;
                    }
                }
                if (found < 0) { bad = true; found = 0; }
                modes[k * nl + i] = found;
                if (k + 1 < N) x[i * N + k] = xv[k + 1];
            }
        }
        if (__any_sync(gm, bad)) return HUGE_VAL;
        L = S.depth;
        LANES(r, S.ng) { orient[r] = 1; agen[r] = 0; }
        LANES(j, S.nv) aflag[j] = 0;
        double bestf = HUGE_VAL;
        int total = 1;
        _Pragma("unroll 1")
        for (int i = 0; i < nl; ++i) total *= 4;
        _Pragma("unroll 1")
        for (int mask = 0; mask < total; ++mask) {
            bool skip = false;
            LANES(i, nl) {
                const int dg = (mask >> (2 * i)) & 3;
                const int r = dg == 0 ? lastm[0] : (dg == 1 ? lastm[1] : (dg == 2 ? lastm[2] : lastm[3]));
                if (r < 0) skip = true;
                else {
                    const double v = xg[(size_t)i * 2 * np1 + np1 + N - 1], uu = ug[i * N + N - 1];
                    modes[(N - 1) * nl + i] = r;
                    x[i * N + N - 1] = ma(i, r) * v + mc(i, r) + mb(i, r) * uu;
                }
            }
            if (__any_sync(gm, skip)) continue;
            __syncwarp(gm);
            LANES(d, L) {
                const int i = d % nl, r = modes[d];
                am[d] = ma(i, r); bm[d] = mb(i, r); cm[d] = mc(i, r);
            }
            __syncwarp(gm);
            double bv; int bid;
            scan(bv, bid, tol, false);
            if (bid == 0x7fffffff) bestf = fmin(bestf, objective());
            __syncwarp(gm);
        }
        return bestf;
    }

    __device__ void finish(double* u_out, double* x_out, double* e_out, int32_t* m_out, double* obj,
                           int32_t* status, int32_t* nodes_out, int32_t* iters_out) {
        const int nl = S.nl, N = S.N, np1 = N + 1;
        const bool ok = own < HUGE_VAL;
        __syncwarp(gm);
        LANES(i, nl) {
            double* xo = x_out + (size_t)i * 2 * np1;
            double* uo = u_out + (size_t)i * N;
            int32_t* mo = m_out + (size_t)i * N;
            if (ok) {
                double p = pvec[2 * i], v = v0[i];
                xo[0] = p; xo[np1] = v;
                _Pragma("unroll 1")
                for (int k = 0; k < N; ++k) {
                    const int r = bmodes[k * nl + i];
                    const double vn = best[i * N + k];
                    uo[k] = (vn - ma(i, r) * v - mc(i, r)) / mb(i, r);
                    mo[k] = r;
                    p = p + v; v = vn;
                    xo[k + 1] = p; xo[np1 + k + 1] = v;
                }
            } else {
                _Pragma("unroll 1")
                for (int k = 0; k < N; ++k) { uo[k] = 0.0; mo[k] = -1; }
                _Pragma("unroll 1")
                for (int k = 0; k <= N; ++k) { xo[k] = 0.0; xo[np1 + k] = 0.0; }
            }
        }
        if (e_out) LANES(e, S.ne) {
            double val = 0.0;
            if (ok) {
                const int ix = S.nel ? S.emap[e] : e;
                if (ix >= 0) val = best[nl * N + ix];
                else {                                              // eliminated copy: zE* = hE + Rz zK
                    const int ee = -1 - ix;
                    val = yel[S.nel + ee] + dot2(S.Rz + (size_t)ee * S.nv, 1, best, S.nv);
                }
            }
            e_out[e] = val;
        }
        if (lane == 0) {
            *obj = ok ? own : HUGE_VAL;
            *status = timeout ? HVP_ST_TIME_LIMIT : limit ? HVP_ST_NODE_LIMIT
                            : (trouble ? HVP_ST_NUMERIC : (ok ? HVP_ST_OPTIMAL : HVP_ST_INFEASIBLE));
            *nodes_out = nodes;
            if (iters_out) *iters_out = iters;
        }
        __syncwarp(gm);
    }
};

}  // namespace

// Adoption, receiving side (lane 0 polls): registers as waiting and returns true once a donor has posted a job in this
// worker's mailbox, false when every started worker is waiting (nothing left anywhere: a busy worker is never waiting).
__device__ __forceinline__ bool pm_wait_job(const PmSplit& sp, int wid, int lane, unsigned gm, int GW) {
    int s = 0;
    if (lane == 0) {
        atomicAdd(sp.ad_count, 1ull);
        __threadfence();
        atomicExch(sp.mail_state + wid, 1);
        for (;;) {
            s = *reinterpret_cast<volatile int*>(sp.mail_state + wid);
            if (s == 2) break;
            const unsigned long long c = *reinterpret_cast<volatile unsigned long long*>(sp.ad_count);
            if (s == 1 && (unsigned)(c >> 32) == (unsigned)c && atomicCAS(sp.mail_state + wid, 1, 0) == 1) { s = 0; break; }
            __nanosleep(400);
        }
        __threadfence();
    }
    return __shfl_sync(gm, s, 0, GW) == 2;
}

template <int GW, bool ONE, bool FX, bool AD = false>
__global__ void __launch_bounds__(128)
pm_miqp_kernel(const __grid_constant__ PmDev S, int64_t batch, const double* __restrict__ x0,
               const double* __restrict__ mass, const double* __restrict__ params,
               const int32_t* __restrict__ fixed_modes, const double* __restrict__ Y, double* __restrict__ u,
               double* __restrict__ x, double* __restrict__ extra, int32_t* __restrict__ modes, double* __restrict__ obj,
               int32_t* __restrict__ status, int32_t* __restrict__ nodes, int32_t* __restrict__ qp_iters,
               unsigned long long* __restrict__ counter, const __grid_constant__ PmSplit sp) {
    extern __shared__ double pm_smem[];
    const int lane = threadIdx.x % GW, wib = threadIdx.x / GW;       // lane within the group, group within the CTA
    const unsigned gm = GW == 32 ? FULL : (0xffffu << (16 * ((threadIdx.x >> 4) & 1)));
    double* base = pm_smem + (size_t)wib * (S.smem_bytes / 8);
    Warp<GW, ONE, FX, AD> W(S, base, lane, gm);
    const size_t sx = (size_t)S.nl * 2 * (S.N + 1), su = (size_t)S.nl * S.N;
    // problems are handed out one at a time (tree sizes vary by orders of magnitude)
    W.sp = &sp;
    const int wid = (int)blockIdx.x * (int)(blockDim.x / GW) + wib;
    bool ad_on = false;
    W.ad_probe = wid;
    if (AD) {
        ad_on = sp.ad && (sp.mode == 2 || sp.mode == 3) && wid < sp.mail_cap;
        if (ad_on && lane == 0) atomicAdd(sp.ad_count, 1ull << 32);
    }
    for (;;) {
        unsigned long long nxt = 0;
        if (lane == 0) nxt = atomicAdd(counter, 1ull);
        const int64_t w = (int64_t)__shfl_sync(gm, nxt, 0, GW);
        int64_t i = w, o = w;                    // problem (input) index, output index
        const int* job = nullptr;                // adopted sub-tree: {flagged index, prefix length, slot, prefix modes}
        W.sub_M = 0; W.sub_D = 0; W.sub_code = 0; W.shared = nullptr; W.stop_nodes = 0;
        if (AD) { W.job_P = 0; W.ad_nwork = 0; }
        W.budget = (sp.mode == 1 && !fixed_modes) ? sp.budget : 0;
        if (sp.mode == 3) {
            // sharded pass: M groups of THIS device per problem; the sub-trees of a problem are dealt to
            // world * M groups over all devices by prefix ordinal; the incumbent slot was seeded by the caller
            if (w >= batch * sp.M) {
                if (!AD || !ad_on || !pm_wait_job(sp, wid, lane, gm, GW)) break;
                job = sp.mail_job + (size_t)wid * sp.mail_stride;
                i = *reinterpret_cast<const volatile int*>(job);
                W.sub_M = 1; W.sub_D = *reinterpret_cast<const volatile int*>(job + 1); W.sub_code = 0;
                o = (int64_t)sp.pool_base + *reinterpret_cast<const volatile int*>(job + 2);
            } else {
                i = w / sp.M;
                W.sub_M = sp.M * sp.world; W.sub_D = sp.D; W.sub_code = (int)(w % sp.M) * sp.world + sp.rank;   // ordinal o -> rank o % world
            }
            W.shared = sp.inc_shared + i;
            W.stop_nodes = sp.budget;
            if (AD && ad_on) { W.fidx = (int)i; W.ad_nwork = min((int)gridDim.x * (int)(blockDim.x / GW), sp.mail_cap); }
        } else if (sp.mode == 2) {
            int nf = *reinterpret_cast<volatile int*>(sp.nflag);
            if (nf > sp.cap) nf = sp.cap;
            int f;
            if (w >= (int64_t)nf * sp.M) {
                if (!AD || !ad_on || !pm_wait_job(sp, wid, lane, gm, GW)) break;
                job = sp.mail_job + (size_t)wid * sp.mail_stride;
                f = *reinterpret_cast<const volatile int*>(job);
                W.sub_M = 1; W.sub_D = *reinterpret_cast<const volatile int*>(job + 1); W.sub_code = 0;
                o = (int64_t)sp.pool_base + *reinterpret_cast<const volatile int*>(job + 2);
            } else {
                f = (int)(w / sp.M);
                W.sub_M = sp.M; W.sub_D = sp.D; W.sub_code = (int)(w % sp.M);
            }
            i = sp.flagged[f];
            W.shared = sp.inc_shared + f;
            if (AD && ad_on) { W.fidx = f; W.ad_nwork = min((int)gridDim.x * (int)(blockDim.x / GW), sp.mail_cap); }
        } else if (w >= batch) break;
        W.prob = i;
        W.setup(x0 + (size_t)i * 2 * S.nl, mass + (size_t)i * S.nl, params + (size_t)i * S.npar,
                fixed_modes ? fixed_modes + su * i : nullptr, Y ? Y + (size_t)S.mw * i : nullptr);
        if (AD && job) {
            // the donor's path becomes the forced prefix; there is an incumbent already, so no first dive
            W.job_P = W.sub_D;
            for (int l = lane; l < W.job_P; l += GW) W.modes[l] = *reinterpret_cast<const volatile int*>(job + 3 + l);
            W.dive = false;
            __syncwarp(gm);
            if (lane == 0) { W.open_level(0); atomicExch(sp.mail_state + wid, 0); }
            __syncwarp(gm);
        }
        W.solve();
        W.finish(u + su * o, x + sx * o, extra ? extra + (size_t)S.ne * o : nullptr, modes + su * o, obj + o,
                 status + o, nodes + o, qp_iters ? qp_iters + o : nullptr);
    }
}

// pass 3 of the tree split: one warp per flagged problem keeps the best of its M sub-results (or the pass-1
// incumbent when no sub-tree improved on it) and adds up the work counters
__global__ void __launch_bounds__(128)
pm_merge_kernel(const __grid_constant__ PmDev S, const __grid_constant__ PmSplit sp, const double* __restrict__ su_,
                const double* __restrict__ sx_, const double* __restrict__ se_, const int32_t* __restrict__ sm_,
                const double* __restrict__ sobj, const int32_t* __restrict__ sst, const int32_t* __restrict__ sno,
                const int32_t* __restrict__ sit, double* __restrict__ u, double* __restrict__ x,
                double* __restrict__ extra, int32_t* __restrict__ modes, double* __restrict__ obj,
                int32_t* __restrict__ status, int32_t* __restrict__ nodes, int32_t* __restrict__ qp_iters) {
    const int lane = threadIdx.x & 31;
    const int f = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    int nf = *sp.nflag;
    if (nf > sp.cap) nf = sp.cap;
    if (f >= nf) return;
    const int i = sp.flagged[f];
    const size_t nu = (size_t)S.nl * S.N, nx = (size_t)S.nl * 2 * (S.N + 1);
    double bestv = obj[i];                       // incumbent of the budgeted pass (+inf if it had none)
    int bw = -1, nsum = 0, isum = 0;
    bool numeric = false, limited = false, timed = false;
    for (int c = 0; c < sp.M; ++c) {
        const size_t w = (size_t)f * sp.M + c;
        if (sobj[w] < bestv) { bestv = sobj[w]; bw = (int)w; }
        nsum += sno[w];
        if (sit) isum += sit[w];
        if (sst[w] == HVP_ST_NUMERIC) numeric = true;
        if (sst[w] == HVP_ST_NODE_LIMIT) limited = true;
        if (sst[w] == HVP_ST_TIME_LIMIT) timed = true;
    }
    if (sp.ad) {
        // results of the adopted sub-trees: pool slots p (at pool_base + p) whose owner is this problem
        int pu = *sp.pool_used;
        if (pu > sp.pool_cap) pu = sp.pool_cap;
        double pv = HUGE_VAL; int pw = -1, pn = 0, pi = 0, pflags = 0;
        for (int p = lane; p < pu; p += 32) {
            if (sp.pool_owner[p] != f) continue;
            const size_t w = (size_t)sp.pool_base + p;
            if (sobj[w] < pv) { pv = sobj[w]; pw = (int)w; }
            pn += sno[w];
            if (sit) pi += sit[w];
            if (sst[w] == HVP_ST_NUMERIC) pflags |= 1;
            if (sst[w] == HVP_ST_NODE_LIMIT) pflags |= 2;
            if (sst[w] == HVP_ST_TIME_LIMIT) pflags |= 4;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, pv, o);
            const int ow = __shfl_xor_sync(0xffffffffu, pw, o);
            if (ov < pv || (ov == pv && ow >= 0 && (pw < 0 || ow < pw))) { pv = ov; pw = ow; }
            pn += __shfl_xor_sync(0xffffffffu, pn, o);
            pi += __shfl_xor_sync(0xffffffffu, pi, o);
            pflags |= __shfl_xor_sync(0xffffffffu, pflags, o);
        }
        if (pw >= 0 && pv < bestv) { bestv = pv; bw = pw; }
        nsum += pn; isum += pi;
        if (pflags & 1) numeric = true;
        if (pflags & 2) limited = true;
        if (pflags & 4) timed = true;
    }
    if (bw >= 0) {
        for (size_t e = lane; e < nu; e += 32) { u[nu * i + e] = su_[nu * bw + e]; modes[nu * i + e] = sm_[nu * bw + e]; }
        for (size_t e = lane; e < nx; e += 32) x[nx * i + e] = sx_[nx * bw + e];
        if (extra) for (int e = lane; e < S.ne; e += 32) extra[(size_t)S.ne * i + e] = se_[(size_t)S.ne * bw + e];
    } else if (sp.mode == 3) {                   // sharded pass: this device's share held nothing better
        for (size_t e = lane; e < nu; e += 32) { u[nu * i + e] = 0.0; modes[nu * i + e] = -1; }
        for (size_t e = lane; e < nx; e += 32) x[nx * i + e] = 0.0;
        if (extra) for (int e = lane; e < S.ne; e += 32) extra[(size_t)S.ne * i + e] = 0.0;
    }
    if (lane == 0) {
        obj[i] = bestv;
        status[i] = timed ? HVP_ST_TIME_LIMIT : limited ? HVP_ST_NODE_LIMIT
                            : (numeric ? HVP_ST_NUMERIC : (bestv < HUGE_VAL ? HVP_ST_OPTIMAL : HVP_ST_INFEASIBLE));
        nodes[i] += nsum;
        if (qp_iters) qp_iters[i] += isum;
    }
}

__global__ void __launch_bounds__(128)
pm_eval_kernel(const __grid_constant__ PmDev S, int64_t batch, const double* __restrict__ x0,
               const double* __restrict__ mass, const double* __restrict__ params,
               const double* __restrict__ xg, const double* __restrict__ ug, double* __restrict__ cost) {
    extern __shared__ double pm_smem[];
    constexpr int GW = 32;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double* base = pm_smem + (size_t)wib * (S.smem_bytes / 8);
    Warp<GW, false> W(S, base, lane, FULL);
    const size_t sx = (size_t)S.nl * 2 * (S.N + 1), su = (size_t)S.nl * S.N;
    for (int64_t i = (int64_t)blockIdx.x * wpb + wib; i < batch; i += (int64_t)gridDim.x * wpb) {
        W.setup(x0 + (size_t)i * 2 * S.nl, mass + (size_t)i * S.nl, params + (size_t)i * S.npar, nullptr);
        const double f = W.eval(xg + sx * i, ug + su * i);
        if (lane == 0) cost[i] = f;
        __syncwarp();
    }
}

// ---- shared-structure multi-RHS product on the FP64 tensor cores --------------------------------
// Every problem of a batch shares the formulation matrices; only pvec = [x0 ; params ; 1] differs.
// So gradient, right-hand sides and the constant term of ALL problems are one dense product
//     Y [batch][mw] = P [batch][kw] . W' [kw][mw]
// issued as mma.sync.m8n8k4 FP64 (DMMA): a warp owns 16 problems (two 8-row tiles of P) and walks the
// output columns 32 at a time; fragments come straight from global memory / L1 (W is a few hundred KB
// at most and shared by every warp; each P row is read once per 32-column sweep from L1).
__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128)
pm_precompute_kernel(const __grid_constant__ PmDev S, int64_t batch, const double* __restrict__ x0,
                     const double* __restrict__ params, double* __restrict__ Y) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t b0 = warp * 16;
    if (b0 >= batch) return;
    const int row = lane >> 2, kk = lane & 3;
    const int nx = 2 * S.nl, npv = S.npv, kw = S.kw, mw = S.mw;
    auto P = [&](int64_t b, int k) -> double {        // element k of problem b's parameter vector
        if (b >= batch || k >= npv) return 0.0;
        if (k < nx) return x0[(size_t)b * nx + k];
        if (k < npv - 1) return params[(size_t)b * S.npar + (k - nx)];
        return 1.0;
    };
    // the 32-column tiles of the output are dealt over gridDim.y: a small batch (the one-vehicle roles of a consensus
    // round: 1024 problems = 16 CTAs) would otherwise walk all of them serially on a handful of SMs (r02i launch list:
    // 113 us per launch, as long as the QPs it prepares)
    const int tiles = (mw + 31) / 32, per_y = (tiles + (int)gridDim.y - 1) / (int)gridDim.y;
    const int n_begin = (int)blockIdx.y * per_y * 32, n_end = min(mw, n_begin + per_y * 32);
    for (int n0 = n_begin; n0 < n_end; n0 += 32) {
        double acc[2][4][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[a][t][0] = acc[a][t][1] = 0.0;
        for (int k0 = 0; k0 < kw; k0 += 4) {
            const double a0 = P(b0 + row, k0 + kk), a1 = P(b0 + 8 + row, k0 + kk);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int n = n0 + t * 8 + row;
                const double bf = (n < mw) ? S.W[(size_t)n * kw + k0 + kk] : 0.0;
                dmma8x8x4(acc[0][t][0], acc[0][t][1], a0, bf);
                dmma8x8x4(acc[1][t][0], acc[1][t][1], a1, bf);
            }
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int64_t b = b0 + a * 8 + row;
            if (b >= batch) continue;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int n = n0 + t * 8 + 2 * kk;
                if (n < mw) {
                    double2 v; v.x = acc[a][t][0]; v.y = acc[a][t][1];
                    *reinterpret_cast<double2*>(Y + (size_t)b * mw + n) = v;
                }
            }
        }
    }
}

cudaError_t launch_pm_precompute(const PmDev& S, int64_t batch, const double* x0, const double* params, double* Y,
                                 cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    const int64_t warps = (batch + 15) / 16;
    const int64_t blocks = (warps + 3) / 4;
    const int tiles = (S.mw + 31) / 32;
    int64_t gy = (8 * 148 + blocks - 1) / blocks;            // aim at ~8 CTAs per SM in total
    if (gy > tiles) gy = tiles;
    if (gy < 1) gy = 1;
    pm_precompute_kernel<<<dim3((unsigned)blocks, (unsigned)gy), 128, 0, stream>>>(S, batch, x0, params, Y);
    return cudaGetLastError();
}

// per-warp shared-memory carve-up
void pm_layout(PmDev& S) {
    const int nv = S.nv, D = S.depth;
    S.ld = nv | 1;
    int o = 0;
    S.o_hinv = o; o += nv * S.ld;
    S.o_ginv = o; o += nv * S.ld;
    S.o_nact = o; o += nv * S.ld;
    S.o_vec = o; o += 12 * nv;
    S.o_cres = o; o += S.nres + S.nlin;
    S.o_bgen = o; o += S.ng;
    S.o_pvec = o; o += S.npv;
    S.o_misc = o; o += 3 * S.nl + (D + 1) + 2 * S.nl * (S.N + 1) + 5 * D + 4 * (D + 1);
    S.smem_doubles = o;
    const int ints = nv + (D + 1) + 3 * D + S.ng + nv + S.ng + 2 * nv;
    S.o_int = o;
    S.smem_bytes = (o * 8 + ints * 4 + 15) / 16 * 16;
}

// Groups (= problems in flight) per CTA and CTAs per SM of pm_miqp_kernel: the kernel is latency bound (r02o ncu: issue slots
// 21 % busy), so what counts is how many groups an SM holds, and that is set by shared memory -- 227 KB minus 1 KB
// per CTA -- and by the register file (~128 registers per thread: 16 warps).  Round 1 always packed 128 threads per
// CTA: 76 KB per CTA at 15 variables = 2 CTAs = 16 groups per SM where 22 fit as eleven one-warp CTAs.
template <int GW>
static void pm_launch_shape(const PmDev& S, int& gpb, int& per_sm) {
    static const int force = getenv("HVP_MPC_GPB") ? atoi(getenv("HVP_MPC_GPB")) : 0;
    const size_t sm_smem = 233472, cta_overhead = 1024;
    int best_g = 0, best_c = 0;
    for (int g = 128 / GW; g >= (GW == 16 ? 2 : 1); g >>= 1) {
        if (force && g != force) continue;
        const size_t smem = (size_t)g * S.smem_bytes;
        if (smem > 227 * 1024) continue;
        int c = (int)(sm_smem / (smem + cta_overhead));
        const int threads = g * GW;
        if (c * threads > 512) c = 512 / threads;          // register file: 16 warps of ~128 registers per thread
        if (c > 32) c = 32;
        if (c < 1) c = 1;
        if (c * g > best_c * best_g) { best_g = g; best_c = c; }
    }
    if (best_g == 0) { best_g = GW == 16 ? 2 : 1; best_c = 1; }
    gpb = best_g; per_sm = best_c;
}

// MaxDynamicSharedMemorySize of pm_miqp_kernel<GW>: grow-only, cached per device, shared by EVERY launch path of the kernel
// (the sharded path used to set its own, smaller value on every call and left the cache of the plain path stale:
// a later launch with more shared memory failed with "invalid argument")
template <int GW, bool ONE, bool FX, bool AD = false>
static cudaError_t pm_miqp_smem_attr(size_t smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    static size_t attr_set[HVP_MAX_DEVICES] = {0};
    if (dev < 0 || dev >= HVP_MAX_DEVICES || smem > attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(pm_miqp_kernel<GW, ONE, FX, AD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < HVP_MAX_DEVICES) attr_set[dev] = smem;
    }
    return cudaSuccess;
}

template <int GW, bool ONE>
static cudaError_t launch_pm_miqp_t(const PmDev& S, int64_t batch, const double* x0, const double* mass,
                                    const double* params, const int32_t* fixed_modes, const double* Y, double* u,
                                    double* x, double* extra, int32_t* modes, double* obj, int32_t* status,
                                    int32_t* nodes, int32_t* qp_iters, unsigned long long* counter,
                                    const PmScratch* sc, cudaStream_t stream) {
    // gpb groups (= problems in flight) per CTA of gpb * GW threads
    int gpb = 0, per_sm = 0;
    pm_launch_shape<GW>(S, gpb, per_sm);
    const size_t smem = (size_t)gpb * S.smem_bytes;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    // function attributes are per DEVICE (one process may drive several: Context(device)): cached per device
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    {
        cudaError_t e = pm_miqp_smem_attr<GW, ONE, false>(smem);
        if (e != cudaSuccess) return e;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = gpb * GW;
    const int64_t full = (int64_t)sms * per_sm;
    int64_t blocks = (batch + gpb - 1) / gpb;
    if (blocks > full) blocks = full;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    const bool split = sc && !fixed_modes && S.max_nodes == 0 && sc->sp.D >= 1;
    PmSplit sp;
    memset(&sp, 0, sizeof sp);
    if (!split) {
        if (fixed_modes) {                   // fixed-sequence QPs (LPs under the 1-norm cost): the instantiation without the search
            e = pm_miqp_smem_attr<GW, ONE, true>(smem);
            if (e != cudaSuccess) return e;
            pm_miqp_kernel<GW, ONE, true><<<(unsigned)blocks, threads, smem, stream>>>(S, batch, x0, mass, params, fixed_modes, Y, u,
                                                                                       x, extra, modes, obj, status, nodes, qp_iters,
                                                                                       counter, sp);
            return cudaGetLastError();
        }
        pm_miqp_kernel<GW, ONE, false><<<(unsigned)blocks, threads, smem, stream>>>(S, batch, x0, mass, params, fixed_modes, Y, u, x,
                                                                        extra, modes, obj, status, nodes, qp_iters, counter, sp);
        return cudaGetLastError();
    }
    // pass 1: every problem under a node budget; heavy trees are appended to the flagged list
    sp = sc->sp;
    sp.mode = 1;
    e = cudaMemsetAsync(sp.nflag, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
    pm_miqp_kernel<GW, ONE, false><<<(unsigned)blocks, threads, smem, stream>>>(S, batch, x0, mass, params, nullptr, Y, u, x, extra,
                                                                    modes, obj, status, nodes, qp_iters, counter, sp);
    // pass 2: M groups per flagged problem, sub-trees by prefix ordinal, incumbent shared through global memory
    e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    sp.mode = 2;
    int64_t b2 = ((int64_t)sp.cap * sp.M + gpb - 1) / gpb;
    if (b2 > full) b2 = full;
    sp.pool_base = sp.cap * sp.M;
    if (sp.ad) {
        // with adoption: the waiting workers' mailboxes and the counters start from zero (one contiguous region)
        e = cudaMemsetAsync(sp.ad_count, 0, 16 + (size_t)sp.mail_cap * sizeof(int), stream);
        if (e != cudaSuccess) return e;
        e = pm_miqp_smem_attr<GW, ONE, false, true>(smem);
        if (e != cudaSuccess) return e;
        pm_miqp_kernel<GW, ONE, false, true><<<(unsigned)b2, threads, smem, stream>>>(S, batch, x0, mass, params, nullptr, Y, sc->u,
                                                                             sc->x, sc->extra, sc->modes, sc->obj, sc->status,
                                                                             sc->nodes, sc->iters, counter, sp);
    } else {
        sp.ad = 0;
        pm_miqp_kernel<GW, ONE, false><<<(unsigned)b2, threads, smem, stream>>>(S, batch, x0, mass, params, nullptr, Y, sc->u, sc->x,
                                                                    sc->extra, sc->modes, sc->obj, sc->status, sc->nodes,
                                                                    sc->iters, counter, sp);
    }
    // pass 3: keep the best sub-result of every flagged problem
    pm_merge_kernel<<<(unsigned)((sp.cap + 3) / 4), 128, 0, stream>>>(S, sp, sc->u, sc->x, sc->extra, sc->modes, sc->obj,
                                                                     sc->status, sc->nodes, sc->iters, u, x, extra,
                                                                     modes, obj, status, nodes, qp_iters);
    return cudaGetLastError();
}

// ---- one device's share of split trees (multi-GPU tree split; the incumbent exchange between the devices is the
// caller's allreduce(min), hvp.h: hvp_mpc_solve_shard_dev) -------------------------------------------------------
__global__ void __launch_bounds__(128)
pm_shard_init_kernel(int64_t batch, const double* __restrict__ incumbent, PmSplit sp, double* __restrict__ obj,
                     int32_t* __restrict__ nodes, int32_t* __restrict__ qp_iters) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *sp.nflag = (int)batch;
    if (i >= batch) return;
    sp.flagged[i] = (int)i;
    sp.inc_shared[i] = dkey(incumbent ? incumbent[i] : HUGE_VAL);
    obj[i] = HUGE_VAL;
    nodes[i] = 0;
    if (qp_iters) qp_iters[i] = 0;
}

template <int GW, bool ONE>
static cudaError_t launch_pm_shard_t(const PmDev& S, int64_t batch, const double* x0, const double* mass,
                                     const double* params, const double* Y, const double* incumbent, double* u,
                                     double* x, double* extra, int32_t* modes, double* obj, int32_t* status,
                                     int32_t* nodes, int32_t* qp_iters, unsigned long long* counter,
                                     const PmScratch* sc, cudaStream_t stream) {
    int gpb = 0, per_sm = 0;
    pm_launch_shape<GW>(S, gpb, per_sm);
    const size_t smem = (size_t)gpb * S.smem_bytes;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = pm_miqp_smem_attr<GW, ONE, false>(smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = gpb * GW;
    const int64_t full = (int64_t)sms * per_sm;
    PmSplit sp = sc->sp;
    sp.mode = 3;
    int64_t blocks = (batch * sp.M + gpb - 1) / gpb;
    if (blocks > full) blocks = full;
    e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    pm_shard_init_kernel<<<(unsigned)((batch + 127) / 128), 128, 0, stream>>>(batch, incumbent, sp, obj, nodes, qp_iters);
    sp.pool_base = (int)(batch * sp.M);
    if (sp.budget > 0) sp.ad = 0;           // a budgeted (probing) wave: an adopted sub-tree would start a budget of its own
    if (sp.ad) {
        e = cudaMemsetAsync(sp.ad_count, 0, 16 + (size_t)sp.mail_cap * sizeof(int), stream);
        if (e != cudaSuccess) return e;
        e = pm_miqp_smem_attr<GW, ONE, false, true>(smem);
        if (e != cudaSuccess) return e;
        pm_miqp_kernel<GW, ONE, false, true><<<(unsigned)blocks, threads, smem, stream>>>(S, batch, x0, mass, params, nullptr, Y, sc->u,
                                                                                 sc->x, sc->extra, sc->modes, sc->obj, sc->status,
                                                                                 sc->nodes, sc->iters, counter, sp);
    } else {
        sp.ad = 0;
        pm_miqp_kernel<GW, ONE, false><<<(unsigned)blocks, threads, smem, stream>>>(S, batch, x0, mass, params, nullptr, Y, sc->u, sc->x,
                                                                        sc->extra, sc->modes, sc->obj, sc->status, sc->nodes,
                                                                        sc->iters, counter, sp);
    }
    pm_merge_kernel<<<(unsigned)((batch + 3) / 4), 128, 0, stream>>>(S, sp, sc->u, sc->x, sc->extra, sc->modes, sc->obj,
                                                                    sc->status, sc->nodes, sc->iters, u, x, extra, modes,
                                                                    obj, status, nodes, qp_iters);
    return cudaGetLastError();
}

cudaError_t launch_pm_shard(const PmDev& S, int64_t batch, const double* x0, const double* mass, const double* params,
                            const double* Y, const double* incumbent, double* u, double* x, double* extra,
                            int32_t* modes, double* obj, int32_t* status, int32_t* nodes, int32_t* qp_iters,
                            unsigned long long* counter, const PmScratch* sc, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
#define HVP_SHARD(GW_, ONE_) launch_pm_shard_t<GW_, ONE_>(S, batch, x0, mass, params, Y, incumbent, u, x, extra, modes, obj, status, nodes, qp_iters, counter, sc, stream)
    if (S.nv <= 16) return S.one_norm ? HVP_SHARD(16, true) : HVP_SHARD(16, false);
    return S.one_norm ? HVP_SHARD(32, true) : HVP_SHARD(32, false);
#undef HVP_SHARD
}

cudaError_t launch_pm_miqp(const PmDev& S, int64_t batch, const double* x0, const double* mass,
                           const double* params, const int32_t* fixed_modes, const double* Y, double* u, double* x,
                           double* extra, int32_t* modes, double* obj, int32_t* status, int32_t* nodes,
                           int32_t* qp_iters, unsigned long long* counter, const PmScratch* sc, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    // problems with at most 16 variables (e.g. the centralized n = 3, N = 5 MIQP) get a 16-lane group: two per warp
    static const int force = getenv("HVP_MPC_GROUP") ? atoi(getenv("HVP_MPC_GROUP")) : 0;
    const bool half = force ? force == 16 : (S.nv <= 16);
#define HVP_MIQP(GW_, ONE_) launch_pm_miqp_t<GW_, ONE_>(S, batch, x0, mass, params, fixed_modes, Y, u, x, extra, modes, obj, status, nodes, qp_iters, counter, sc, stream)
    if (half) return S.one_norm ? HVP_MIQP(16, true) : HVP_MIQP(16, false);
    return S.one_norm ? HVP_MIQP(32, true) : HVP_MIQP(32, false);
#undef HVP_MIQP
}

cudaError_t launch_pm_eval(const PmDev& S, int64_t batch, const double* x0, const double* mass,
                           const double* params, const double* xg, const double* ug, double* cost,
                           cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    int wpb = 4;
    while (wpb > 1 && (size_t)wpb * S.smem_bytes > 200 * 1024) wpb >>= 1;
    const size_t smem = (size_t)wpb * S.smem_bytes;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static size_t attr_set[HVP_MAX_DEVICES] = {0};
    if (dev < 0 || dev >= HVP_MAX_DEVICES || smem > attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(pm_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < HVP_MAX_DEVICES) attr_set[dev] = smem;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t blocks = (batch + wpb - 1) / wpb;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    pm_eval_kernel<<<(unsigned)blocks, wpb * 32, smem, stream>>>(S, batch, x0, mass, params, xg, ug, cost);
    return cudaGetLastError();
}

}  // namespace hvp
