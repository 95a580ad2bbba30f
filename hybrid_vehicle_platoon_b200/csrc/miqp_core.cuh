// miqp_core.cuh -- per-vehicle local hybrid-MPC MIQP, solved by one GPU thread.
//
// Problem: LocalMpcMld of the reference (fleet_decent_mld.py:61-208, fleet_seq_mld.py:63-219) on
// the PWA-gear model (models.py:397-492) in MLD form (dmpcpwa MpcMld).
//
// B200-first formulation (NOT the reference's big-M model):
//   * variables are the VELOCITIES v_1..v_N (positions are prefix sums, inputs are the affine
//     map u_k = (v_{k+1} - a_k v_k - c_k)/b_k of the active region), so region / state / accel /
//     input rows are 1- or 2-sparse and the safe-distance rows are prefix sums;
//   * the slack variables are eliminated exactly: w*s with s >= max(0, g(v)) is an L1 penalty,
//     i.e. a row whose multiplier is bounded by w (bounded-dual Goldfarb-Idnani);
//   * branch-and-bound over the region sequence, depth first, child nearest to the relaxed
//     velocity first; a node fixes the regions of stages 0..L-1 and relaxes stages L..N-1
//     (no input cost, no input bounds, only accel + state box) -- a valid lower bound because
//     every dropped term is non-negative and every dropped row only enlarges the feasible set.
// Everything a thread needs lives in registers / L1-resident local memory (~1.7 kB at N=6).
//
// The file is host+device so that tests can compile it with g++ and check the algorithm
// against the oracle without a GPU (tests/host_harness); the product only ever runs it inside
// the CUDA kernels of local_miqp.cu.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define HVP_HD __host__ __device__ __forceinline__
#define HVP_HDN __host__ __device__ __noinline__
#else
#define HVP_HD inline
#define HVP_HDN
#endif

namespace hvp {

constexpr int NREG = 7;

// Constants shared by every problem of a batch (filled on the host, see vehicle_model.h).
struct LocalParams {
    int N;
    int max_nodes;
    double d0, t0, tight;
    double qxp, qxv, qu, w;              // Params.Q_x, Q_u, w (common_controller_params.py:14-23)
    double a_acc, a_dec, d_safe;
    double vmin, vmax, pmin, pmax;       // D,E rows (models.py:474-475)
    double umin, umax;                   // F,G rows (models.py:476-479)
    double edge[NREG + 1];               // region r is edge[r] <= v <= edge[r+1] (models.py:414-444)
    double bgear[NREG];                  // traction gain of region r before /m (models.py:458-466)
    double c1, c2, dfr, mug;             // PWA friction (models.py:276-282), mu*g
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
#define HVP_ST_NUMERIC 12

// constraint types of the velocity-space node QP (all written as  n'x <= rhs)
enum : int { T_UB = 0, T_LB, T_ACC, T_DEC, T_UHI, T_ULO, T_PHI, T_PLO, T_SF, T_SB, T_COUNT };

template <int NMAX>
struct LocalSolver {
    // ---- problem data -------------------------------------------------------------------
    const LocalParams* P;
    int N, flags;
    double p0, v0, pc;                 // pc = p_1 = p0 + v0 (ts = 1)
    double ra[NREG], rb[NREG], rc[NREG];  // region dynamics for this vehicle's mass
    // mode-independent tracking quadratic in x = (v_1..v_N):  1/2 x'Ht x + gt'x + ct
    double Hoff[NMAX];                 // Ht[i][j] for i<j depends only on j: Hoff[j]
    double Hdiag[NMAX];
    double gt[NMAX];
    double ct;
    // prefix-sum rows:  PS_j = x_0+..+x_{j-1} (= p_{j+1} - pc),  j = 1..N-1
    double sf_rhs[NMAX], sb_rhs[NMAX]; // soft rows  PS_j <= sf_rhs[j],  PS_j >= sb_rhs[j]
    bool has_sf, has_sb;
    // ---- node data ----------------------------------------------------------------------
    int L;                             // stages 0..L-1 have fixed regions
    int modes[NMAX];
    double lb[NMAX], ub[NMAX];         // merged simple bounds on x_j = v_{j+1}
    // ---- QP work space ------------------------------------------------------------------
    double H[NMAX][NMAX], Hinv[NMAX][NMAX];
    double g[NMAX], c0;
    double Nact[NMAX][NMAX], G[NMAX][NMAX], GL[NMAX][NMAX];
    double x[NMAX], lam[NMAX], r[NMAX], d[NMAX], z[NMAX], yp[NMAX], nrow[NMAX], wv[NMAX];
    int act[NMAX];
    int q;
    uint32_t amask[T_COUNT];           // active-set membership, bit j of type t
    uint32_t satf, satb;               // soft rows currently in "violated" orientation
    int iters;

    // -------------------------------------------------------------------------------------
    HVP_HD void setup(const LocalParams* P_, int flags_, double mass, const double* x0,
                      const double* xf, const double* xb, const double* xl) {
        P = P_; N = P->N; flags = flags_;
        p0 = x0[0]; v0 = x0[1]; pc = p0 + v0;
        for (int rg = 0; rg < NREG; ++rg) {      // models.py:447-473 + forward_euler (ts = 1)
            double fr = (rg < 4) ? P->c1 : P->c2;
            ra[rg] = 1.0 + (-(fr) / mass);
            rb[rg] = P->bgear[rg] / mass;
            rc[rg] = (rg < 4) ? (-P->mug) : (-P->mug - P->dfr / mass);
        }
        const bool is_front = flags & 1, is_leader = flags & 2, is_trailer = flags & 4;
        // tracking terms  wp*(p_k + tau*v_k - Pk)^2 + wvv*(v_k - Vk)^2
        //   front  (fleet_decent_mld.py:110-121): tau = t0, Pk = pf_k - d0,           Vk = vf_k
        //   back   (:122-133):                    tau = 0,  Pk = pb_k + t0*vb_k + d0, Vk = vb_k
        //   leader (:134-141):                    tau = 0,  Pk = pl_k,                Vk = vl_k
        const bool tf = !is_front && !is_leader, tb = !is_trailer && !is_leader, tl = is_leader;
        const double wp = P->qxp, wvv = P->qxv, t0 = P->t0, d0 = P->d0;
        const int np1 = N + 1;
        double nterm = (tf ? 1.0 : 0.0) + (tb ? 1.0 : 0.0) + (tl ? 1.0 : 0.0);
        double tau_sum = tf ? t0 : 0.0, tau2_sum = tf ? t0 * t0 : 0.0;
        for (int j = 0; j < N; ++j) {            // x_j = v_{j+1}; stage k = j+1
            // rows m_k (k=1..N) have 1 at columns < k-1.. in x-indexing: 1 for col < k-1, tau at col k-1
            // H[i][j] (i<j): sum over k-1 > j of 1*1 (N-1-j rows) + (k-1 == j): 1*tau
            Hoff[j] = 2.0 * wp * (nterm * (double)(N - 1 - j) + tau_sum);
            Hdiag[j] = 2.0 * wp * (nterm * (double)(N - 1 - j) + tau2_sum) + 2.0 * wvv * nterm;
            gt[j] = 0.0;
        }
        ct = 0.0;
        // rho_k = pc - Pk for k>=1 ; g[j] = 2wp*(sum_{k-1>j} rho_k + tau*rho_{j+1}) - 2wv*V_{j+1}
        auto add_term = [&](const double* ref, double tau, int kind) {
            double suffix = 0.0;
            for (int k = N; k >= 1; --k) {
                double pk = ref[k], vk = ref[np1 + k];
                double Pk = (kind == 0) ? (pk - d0) : (kind == 1) ? (pk + t0 * vk + d0) : pk;
                double rho = pc - Pk;
                int j = k - 1;
                gt[j] += 2.0 * wp * (suffix + tau * rho) - 2.0 * wvv * vk;
                ct += wp * rho * rho + wvv * vk * vk;
                suffix += rho;
            }
            double pk = ref[0], vk = ref[np1];
            double Pk = (kind == 0) ? (pk - d0) : (kind == 1) ? (pk + t0 * vk + d0) : pk;
            double e0 = p0 + tau * v0 - Pk, e1 = v0 - vk;
            ct += wp * e0 * e0 + wvv * e1 * e1;    // k = 0 terms are constants but count in objVal
        };
        if (tf) add_term(xf, t0, 0);
        if (tb) add_term(xb, 0.0, 1);
        if (tl) add_term(xl, 0.0, 2);
        // soft safe-distance rows (fleet_decent_mld.py:191-208); k = 0,1 are constants
        has_sf = !is_front; has_sb = !is_trailer;
        const double w = P->w, ds = P->d_safe;
        if (has_sf) {
            double s0 = p0 - (xf[0] - ds), s1 = pc - (xf[1] - ds);
            ct += w * (s0 > 0 ? s0 : 0.0) + w * (s1 > 0 ? s1 : 0.0);
            for (int j = 1; j < N; ++j) sf_rhs[j] = xf[j + 1] - ds - pc;
        }
        if (has_sb) {
            double s0 = (xb[0] + ds) - p0, s1 = (xb[1] + ds) - pc;
            ct += w * (s0 > 0 ? s0 : 0.0) + w * (s1 > 0 ? s1 : 0.0);
            for (int j = 1; j < N; ++j) sb_rhs[j] = xb[j + 1] + ds - pc;
        }
    }

    // -------------------------------------------------------------------------------------
    // small dense helpers
    HVP_HD static bool chol(int n, double (*A)[NMAX]) {   // in place, lower
        for (int j = 0; j < n; ++j) {
            double dd = A[j][j];
            for (int k = 0; k < j; ++k) dd -= A[j][k] * A[j][k];
            if (!(dd > 0.0)) return false;
            dd = sqrt(dd);
            A[j][j] = dd;
            double inv = 1.0 / dd;
            for (int i = j + 1; i < n; ++i) {
                double s = A[i][j];
                for (int k = 0; k < j; ++k) s -= A[i][k] * A[j][k];
                A[i][j] = s * inv;
            }
        }
        return true;
    }
    HVP_HD static void chol_solve(int n, const double (*Lm)[NMAX], double* y) {
        for (int i = 0; i < n; ++i) {
            double s = y[i];
            for (int k = 0; k < i; ++k) s -= Lm[i][k] * y[k];
            y[i] = s / Lm[i][i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = y[i];
            for (int k = i + 1; k < n; ++k) s -= Lm[k][i] * y[k];
            y[i] = s / Lm[i][i];
        }
    }

    // dense normal + rhs of constraint (t, j) in its current orientation
    HVP_HD void make_row(int t, int j, double* nr, double& rhs) const {
        for (int i = 0; i < N; ++i) nr[i] = 0.0;
        switch (t) {
            case T_UB: nr[j] = 1.0; rhs = ub[j]; break;
            case T_LB: nr[j] = -1.0; rhs = -lb[j]; break;
            case T_ACC: nr[j] = 1.0; nr[j - 1] = -1.0; rhs = P->a_acc - j * P->tight; break;
            case T_DEC: nr[j] = -1.0; nr[j - 1] = 1.0; rhs = -(P->a_dec + j * P->tight); break;
            case T_UHI: { int rg = modes[j]; nr[j] = 1.0; nr[j - 1] = -ra[rg]; rhs = rc[rg] + rb[rg] * P->umax; } break;
            case T_ULO: { int rg = modes[j]; nr[j] = -1.0; nr[j - 1] = ra[rg]; rhs = -(rc[rg] + rb[rg] * P->umin); } break;
            case T_PHI: for (int i = 0; i < j; ++i) nr[i] = 1.0; rhs = P->pmax - pc; break;
            case T_PLO: for (int i = 0; i < j; ++i) nr[i] = -1.0; rhs = -(P->pmin - pc); break;
            case T_SF: {
                double o = ((satf >> j) & 1u) ? -1.0 : 1.0;
                for (int i = 0; i < j; ++i) nr[i] = o;
                rhs = o * sf_rhs[j];
            } break;
            default: {  // T_SB:  -PS_j <= -sb_rhs[j]
                double o = ((satb >> j) & 1u) ? -1.0 : 1.0;
                for (int i = 0; i < j; ++i) nr[i] = -o;
                rhs = -o * sb_rhs[j];
            } break;
        }
    }
    HVP_HD bool is_soft(int t) const { return t >= T_SF; }

    // -------------------------------------------------------------------------------------
    // Build + solve the QP of the current node (modes[0..L-1] fixed). Returns 0 optimal,
    // 1 infeasible, 2 numeric/iteration trouble.  On success x[] and *obj are set.
    HVP_HDN int solve_node(double* obj) {
        const double tol = 1e-9;
        // ---- Hessian / gradient ----
        for (int i = 0; i < N; ++i) {
            for (int j = 0; j < N; ++j) H[i][j] = (i == j) ? Hdiag[i] : Hoff[i > j ? i : j];
            g[i] = gt[i];
        }
        c0 = ct;
        const double qu = P->qu;
        {   // stage 0 (x_{-1} = v0 is a constant): u_0 = (x_0 - (a v0 + c))/b
            int rg = modes[0];
            double ib = 1.0 / rb[rg], k0 = (ra[rg] * v0 + rc[rg]) * ib;
            H[0][0] += 2.0 * qu * ib * ib;
            g[0] += -2.0 * qu * k0 * ib;
            c0 += qu * k0 * k0;
        }
        for (int k = 1; k < L; ++k) {   // u_k = (x_k - a x_{k-1} - c)/b
            int rg = modes[k];
            double ib = 1.0 / rb[rg], ea = -ra[rg] * ib, kc = -rc[rg] * ib;
            H[k][k] += 2.0 * qu * ib * ib;
            H[k - 1][k - 1] += 2.0 * qu * ea * ea;
            H[k][k - 1] += 2.0 * qu * ea * ib;
            H[k - 1][k] += 2.0 * qu * ea * ib;
            g[k] += 2.0 * qu * kc * ib;
            g[k - 1] += 2.0 * qu * kc * ea;
            c0 += qu * kc * kc;
        }
        // ---- H^-1 via Cholesky ----
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) GL[i][j] = H[i][j];
        if (!chol(N, GL)) return 2;
        for (int cidx = 0; cidx < N; ++cidx) {
            for (int i = 0; i < N; ++i) yp[i] = (i == cidx) ? 1.0 : 0.0;
            chol_solve(N, GL, yp);
            for (int i = 0; i < N; ++i) Hinv[i][cidx] = yp[i];
        }
        for (int i = 0; i < N; ++i) {
            double s = 0.0;
            for (int j = 0; j < N; ++j) s -= Hinv[i][j] * g[j];
            x[i] = s;
        }
        q = 0; satf = 0; satb = 0;
        for (int t = 0; t < T_COUNT; ++t) amask[t] = 0;
        const int maxit = 40 * N + 60;
        int it = 0;
        const double w = P->w;

        for (;;) {
            // ---- most violated row (structured scan, prefix sums on the fly) ----
            double best = tol; int pt = -1, pj = 0;
            double PS = 0.0;
#define HVP_CAND(T, J, S)                                                          \
    {                                                                              \
        double s__ = (S);                                                          \
        if (s__ > best && !((amask[T] >> (J)) & 1u)) { best = s__; pt = (T); pj = (J); } \
    }
            for (int j = 0; j < N; ++j) {
                double xv = x[j];
                HVP_CAND(T_UB, j, xv - ub[j]);
                HVP_CAND(T_LB, j, lb[j] - xv);
                if (j >= 1) {
                    double dv = xv - x[j - 1];
                    HVP_CAND(T_ACC, j, dv - (P->a_acc - j * P->tight));
                    HVP_CAND(T_DEC, j, (P->a_dec + j * P->tight) - dv);
                    if (j < L) {
                        int rg = modes[j];
                        double du = xv - ra[rg] * x[j - 1] - rc[rg];
                        HVP_CAND(T_UHI, j, du - rb[rg] * P->umax);
                        HVP_CAND(T_ULO, j, rb[rg] * P->umin - du);
                    }
                    HVP_CAND(T_PHI, j, PS - (P->pmax - pc));
                    HVP_CAND(T_PLO, j, (P->pmin - pc) - PS);
                    if (has_sf) {
                        double s = PS - sf_rhs[j];
                        HVP_CAND(T_SF, j, ((satf >> j) & 1u) ? -s : s);
                    }
                    if (has_sb) {
                        double s = sb_rhs[j] - PS;
                        HVP_CAND(T_SB, j, ((satb >> j) & 1u) ? -s : s);
                    }
                }
                PS += xv;
            }
#undef HVP_CAND
            if (pt < 0) break;
            double lam_p = 0.0;
            for (;;) {
                if (++it > maxit) { iters += it; return 2; }
                double rhs;
                make_row(pt, pj, nrow, rhs);
                double cp = -rhs;
                for (int i = 0; i < N; ++i) cp += nrow[i] * x[i];
                if (cp <= tol) break;
                // yp = H^-1 n_p
                double nHn = 0.0;
                for (int i = 0; i < N; ++i) {
                    double s = 0.0;
                    for (int j = 0; j < N; ++j) s += Hinv[i][j] * nrow[j];
                    yp[i] = s;
                }
                for (int i = 0; i < N; ++i) nHn += nrow[i] * yp[i];
                // d = Nact yp ; r = G^-1 d ; w = n_p - Nact' r ; z = H^-1 w
                for (int a = 0; a < q; ++a) {
                    double s = 0.0;
                    for (int i = 0; i < N; ++i) s += Nact[a][i] * yp[i];
                    d[a] = s; r[a] = s;
                }
                if (q > 0) {
                    for (int a = 0; a < q; ++a)
                        for (int b2 = 0; b2 <= a; ++b2) GL[a][b2] = G[a][b2];
                    if (!chol(q, GL)) { iters += it; return 2; }
                    chol_solve(q, GL, r);
                }
                for (int i = 0; i < N; ++i) {
                    double s = nrow[i];
                    for (int a = 0; a < q; ++a) s -= r[a] * Nact[a][i];
                    wv[i] = s;
                }
                double nz = 0.0;
                for (int i = 0; i < N; ++i) {
                    double s = 0.0;
                    for (int j = 0; j < N; ++j) s += Hinv[i][j] * wv[j];
                    z[i] = s;
                    nz += nrow[i] * s;
                }
                const bool dependent = (q == N) || !(nz > 1e-11 * nHn);
                const double INF = HUGE_VAL;
                double t2 = dependent ? INF : cp / nz;
                double t1 = INF, t3 = INF;
                int k1 = -1, k3 = -1;
                for (int a = 0; a < q; ++a) {
                    if (r[a] > 1e-14) {
                        double t = lam[a] / r[a];
                        if (t < t1) { t1 = t; k1 = a; }
                    } else if (r[a] < -1e-14 && is_soft(act[a] / NMAX)) {
                        double t = (w - lam[a]) / (-r[a]);
                        if (t < t3) { t3 = t; k3 = a; }
                    }
                }
                double t3p = is_soft(pt) ? (w - lam_p) : INF;
                double t = fmin(fmin(t1, t2), fmin(t3, t3p));
                if (!(t < INF)) { iters += it; return 1; }   // infeasible node
                if (!dependent)
                    for (int i = 0; i < N; ++i) x[i] -= t * z[i];
                for (int a = 0; a < q; ++a) lam[a] -= t * r[a];
                lam_p += t;
                if (t == t2) {                      // full step: p joins the active set
                    for (int i = 0; i < N; ++i) Nact[q][i] = nrow[i];
                    for (int a = 0; a < q; ++a) { G[q][a] = d[a]; G[a][q] = d[a]; }
                    G[q][q] = nHn;
                    act[q] = pt * NMAX + pj; lam[q] = lam_p;
                    amask[pt] |= (1u << pj);
                    ++q;
                    break;
                }
                if (t == t3p) {                     // soft p saturates before becoming feasible
                    if (pt == T_SF) satf ^= (1u << pj); else satb ^= (1u << pj);
                    break;
                }
                int drop;
                if (t == t1) drop = k1;
                else {                               // active soft row saturates: flip + drop
                    drop = k3;
                    int tt = act[drop] / NMAX, jj = act[drop] % NMAX;
                    if (tt == T_SF) satf ^= (1u << jj); else satb ^= (1u << jj);
                }
                {
                    int tt = act[drop] / NMAX, jj = act[drop] % NMAX;
                    amask[tt] &= ~(1u << jj);
                }
                for (int a = drop; a + 1 < q; ++a) {
                    act[a] = act[a + 1]; lam[a] = lam[a + 1];
                    for (int i = 0; i < N; ++i) Nact[a][i] = Nact[a + 1][i];
                }
                for (int a = 0; a < q; ++a)
                    for (int b2 = drop; b2 + 1 < q; ++b2) G[a][b2] = G[a][b2 + 1];
                for (int a = drop; a + 1 < q; ++a)
                    for (int b2 = 0; b2 < q; ++b2) G[a][b2] = G[a + 1][b2];
                --q;
            }
        }
        iters += it;
        // ---- objective at x (original H, g, c0 + L1 penalties) ----
        double f = c0;
        for (int i = 0; i < N; ++i) {
            double s = 0.0;
            for (int j = 0; j < N; ++j) s += H[i][j] * x[j];
            f += x[i] * (0.5 * s + g[i]);
        }
        if (has_sf || has_sb) {
            double PS = 0.0;
            for (int j = 0; j < N; ++j) {
                if (j >= 1) {
                    if (has_sf) { double s = PS - sf_rhs[j]; if (s > 0) f += w * s; }
                    if (has_sb) { double s = sb_rhs[j] - PS; if (s > 0) f += w * s; }
                }
                PS += x[j];
            }
        }
        *obj = f;
        return 0;
    }

    // -------------------------------------------------------------------------------------
    // Branch and bound.  Outputs: u[N], xtraj[(2)(N+1)] row-major, mode_out[N].
    HVP_HD LocalResult solve(double* u_out, double* x_out, int32_t* mode_out) {
        LocalResult R;
        R.obj = HUGE_VAL; R.status = HVP_ST_INFEASIBLE; R.nodes = 0; R.qp_iters = 0;
        iters = 0;
        double inc = HUGE_VAL;
        double best_x[NMAX];
        int best_modes[NMAX];
        uint32_t cand[NMAX + 1];
        double xstar[NMAX + 1];
        double rlo[NMAX + 1], rhi[NMAX + 1];   // reachable velocity interval of v_level
        const double eps = 1e-9;
        bool trouble = false, limit = false;

        int lev = 0;
        rlo[0] = v0; rhi[0] = v0; xstar[0] = v0;
        cand[0] = 0;
        for (int rg = 0; rg < NREG; ++rg)
            if (v0 >= P->edge[rg] && v0 <= P->edge[rg + 1]) cand[0] |= (1u << rg);

        for (;;) {
            if (cand[lev] == 0) {
                if (lev == 0) break;
                --lev;
                continue;
            }
            // candidate region nearest to the parent's relaxed velocity
            int rg = -1; double bd = HUGE_VAL;
            for (int c = 0; c < NREG; ++c) {
                if (!((cand[lev] >> c) & 1u)) continue;
                double lo = P->edge[c], hi = P->edge[c + 1];
                double dist = xstar[lev] < lo ? lo - xstar[lev] : (xstar[lev] > hi ? xstar[lev] - hi : 0.0);
                if (dist < bd) { bd = dist; rg = c; }
            }
            cand[lev] &= ~(1u << rg);
            modes[lev] = rg;
            // reachable interval of v_{lev+1} through region rg (provable pruning only)
            double jlo = fmax(rlo[lev], P->edge[rg]), jhi = fmin(rhi[lev], P->edge[rg + 1]);
            if (jlo > jhi + eps) continue;
            double nlo = fmax(ra[rg] * jlo + rc[rg] + rb[rg] * P->umin, jlo + P->a_dec + lev * P->tight);
            double nhi = fmin(ra[rg] * jhi + rc[rg] + rb[rg] * P->umax, jhi + P->a_acc - lev * P->tight);
            nlo = fmax(nlo, P->vmin); nhi = fmin(nhi, P->vmax);
            if (nlo > nhi + eps) continue;
            rlo[lev + 1] = nlo - eps; rhi[lev + 1] = nhi + eps;
            L = lev + 1;
            // merged simple bounds: state box (k>=1), region of fixed stages k=1..L-1 (x_{k-1}),
            // stage-0 accel and input rows (both bounds on x_0 = v_1)
            for (int j = 0; j < N; ++j) { lb[j] = P->vmin; ub[j] = P->vmax; }
            for (int k = 1; k < L; ++k) {
                lb[k - 1] = fmax(lb[k - 1], P->edge[modes[k]]);
                ub[k - 1] = fmin(ub[k - 1], P->edge[modes[k] + 1]);
            }
            {
                int r0 = modes[0];
                lb[0] = fmax(lb[0], fmax(v0 + P->a_dec, ra[r0] * v0 + rc[r0] + rb[r0] * P->umin));
                ub[0] = fmin(ub[0], fmin(v0 + P->a_acc, ra[r0] * v0 + rc[r0] + rb[r0] * P->umax));
            }
            // p_1 = pc is a constant: its state rows are a feasibility check
            if (pc > P->pmax + eps || pc < P->pmin - eps) continue;
            double obj;
            int st = solve_node(&obj);
            ++R.nodes;
            if (st == 2) { trouble = true; continue; }
            if (st == 1) continue;
            if (inc < HUGE_VAL && !(obj < inc - 1e-9 * fmax(1.0, fabs(inc)))) continue;   // bound
            if (L == N) {                                   // leaf: new incumbent
                inc = obj;
                for (int j = 0; j < N; ++j) { best_x[j] = x[j]; best_modes[j] = modes[j]; }
                continue;
            }
            if (P->max_nodes > 0 && R.nodes >= P->max_nodes) { limit = true; break; }
            ++lev;
            xstar[lev] = x[lev - 1];                        // relaxed v_lev
            cand[lev] = 0;
            for (int c = 0; c < NREG; ++c)
                if (P->edge[c] <= rhi[lev] && P->edge[c + 1] >= rlo[lev]) cand[lev] |= (1u << c);
        }
        R.qp_iters = iters;
        if (inc < HUGE_VAL) {
            R.obj = inc;
            R.status = limit ? HVP_ST_NODE_LIMIT : (trouble ? HVP_ST_NUMERIC : HVP_ST_OPTIMAL);
            const int np1 = N + 1;
            double p = p0, v = v0;
            x_out[0] = p; x_out[np1] = v;
            for (int k = 0; k < N; ++k) {
                int rg = best_modes[k];
                double vn = best_x[k];
                u_out[k] = (vn - ra[rg] * v - rc[rg]) / rb[rg];
                mode_out[k] = rg;
                p = p + v; v = vn;
                x_out[k + 1] = p; x_out[np1 + k + 1] = v;
            }
        } else {
            R.status = limit ? HVP_ST_NODE_LIMIT : (trouble ? HVP_ST_NUMERIC : HVP_ST_INFEASIBLE);
            const int np1 = N + 1;
            for (int k = 0; k < N; ++k) { u_out[k] = 0.0; mode_out[k] = -1; }
            for (int k = 0; k <= N; ++k) { x_out[k] = 0.0; x_out[np1 + k] = 0.0; }
        }
        return R;
    }
};

}  // namespace hvp
