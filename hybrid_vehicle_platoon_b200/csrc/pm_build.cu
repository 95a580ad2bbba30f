// pm_build.cu -- host side of the compiled-MPC entry points (include/hvp.h: hvp_mpc_*).
//
// hvp_mpc_create() plays the role of the reference's model construction (MpcMldCent.__init__,
// LocalMpc.__init__, ...: one Gurobi model per controller, built once): it states the cost and the
// coupling rows of a formulation as affine expressions of the decision vector z and the per-solve
// parameter vector pvec (pm_types.h), assembles the dense matrices the kernel needs (H0, H0^-1,
// residual maps, coupling rows) and keeps them on the device.  hvp_mpc_solve_* then only moves
// (x0, masses, parameters) in and the solution out.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "../../include/hvp.h"
#include "hvp_internal.h"
#include "vehicle_model.h"

using namespace hvp;

#define fail hvp_fail
#define CUDA_TRY(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(-100 - (int)e__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                   \
    } while (0)

// Adoption scratch of a sub-tree pass (PmSplit::ad, pm_kernel.cu): result slots of the adopted sub-trees (the POOL,
// appended to the work items' slots), one mailbox per worker, and the counters -- in ONE region with the mailbox states,
// so that a single memset arms a launch.
static const size_t PM_POOL = 32768, PM_MAIL_CAP = 20480;
static int pm_env_adopt() { static const int v = getenv("HVP_MPC_ADOPT") ? atoi(getenv("HVP_MPC_ADOPT")) : 1; return v; }
static int pm_env_adopt_free() { static const int v = getenv("HVP_MPC_ADOPT_FREE") ? atoi(getenv("HVP_MPC_ADOPT_FREE")) : 2; return v; }
static size_t pm_adopt_bytes(int depth) {
    return (16 + PM_MAIL_CAP * 4 + 256) + (PM_POOL * 4 + 256) + (PM_MAIL_CAP * ((size_t)depth + 3) * 4 + 256);
}
template <class Take>
static void pm_adopt_take(hvp::PmSplit& sp, Take&& take, int depth) {
    char* z = take(16 + PM_MAIL_CAP * 4);            // [ad_count 8][pool_used 4][pad 4][mail_state]
    sp.ad_count = (unsigned long long*)z; sp.pool_used = (int*)(z + 8); sp.mail_state = (int*)(z + 16);
    sp.pool_owner = (int*)take(PM_POOL * 4);
    sp.mail_stride = depth + 3;
    sp.mail_job = (int*)take(PM_MAIL_CAP * (size_t)sp.mail_stride * 4);
    sp.pool_cap = (int)PM_POOL; sp.mail_cap = (int)PM_MAIL_CAP;
    sp.ad = pm_env_adopt(); sp.ad_free = pm_env_adopt_free();
}

struct hvp_mpc {
    // Calls on one handle are serialised: the scratch buffers (ybuf, split / shard scratch, work counter) belong to the
    // handle, and a call both (re)allocates them and enqueues a launch SEQUENCE that must not interleave with another
    // thread's (dist.ThreadRanks plays several ranks on one handle).  Concurrent calls must also use the same stream:
    // the device-side scratch is shared.
    std::mutex mu;
    hvp_ctx* ctx;
    hvp_mpc_desc desc;
    PmDev S;
    std::vector<void*> dev;     // device allocations owned by the handle
    unsigned long long* counter; // work-distribution counter of the kernel
    double* x0buf = nullptr;     // scratch of hvp_mpc_eval_dev
    size_t x0cap = 0;
    double* ybuf = nullptr;      // output of the tensor-core precompute, [batch][mw]
    size_t ycap = 0;
    PmScratch scratch;           // tree-split scratch (pm_types.h); scratch.sp.D == 0: disabled
    void* scratch_mem = nullptr;
    size_t scratch_cap = 0;      // flagged-list capacity the scratch was sized for
    int split_D = 0;
    PmScratch shard;             // scratch of hvp_mpc_solve_shard_dev: batch x groups work items
    void* shard_mem = nullptr;
    size_t shard_items = 0;
};

// hvp_mpc_solve_host / hvp_mpc_eval_host call the *_dev entry points on the same handle: the lock is taken once per thread
static thread_local int hvp_mpc_locked_by_this_thread = 0;
struct HvpReentry {
    HvpReentry() { ++hvp_mpc_locked_by_this_thread; }
    ~HvpReentry() { --hvp_mpc_locked_by_this_thread; }
};

namespace {

// affine expression  z . vars + p . pvec
struct Aff {
    std::vector<double> z, p;
    Aff(int nv, int npv) : z(nv, 0.0), p(npv, 0.0) {}
    Aff& operator+=(const Aff& o) {
        for (size_t i = 0; i < z.size(); ++i) z[i] += o.z[i];
        for (size_t i = 0; i < p.size(); ++i) p[i] += o.p[i];
        return *this;
    }
    bool has_z() const {
        for (double v : z) if (v != 0.0) return true;
        return false;
    }
};
Aff operator+(Aff a, const Aff& b) { a += b; return a; }
Aff operator*(double c, Aff a) {
    for (double& v : a.z) v *= c;
    for (double& v : a.p) v *= c;
    return a;
}
Aff operator-(Aff a, const Aff& b) { a += (-1.0) * b; return a; }

struct Builder {
    int nl, N, ne, nv, npar, npv;
    std::vector<Aff> res; std::vector<double> wres;
    std::vector<Aff> la, lb;                       // cost += la(pvec) * lb(z, pvec)
    std::vector<Aff> rows; std::vector<double> wmax;   // e <= 0
    bool one_norm = false;                             // cost terms are w |e| instead of w e^2 (cent_mld.py:58-61)
    Builder(int nl_, int N_, int ne_, int npar_)
        : nl(nl_), N(N_), ne(ne_), nv(nl_ * N_ + ne_), npar(npar_), npv(2 * nl_ + npar_ + 1) {}
    Aff zero() const { return Aff(nv, npv); }
    Aff P(int i, int k) const {                    // position of local vehicle i at stage k (ts = 1)
        Aff a = zero();
        a.p[2 * i] = 1.0;
        if (k >= 1) {
            a.p[2 * i + 1] = 1.0;
            for (int j = 1; j < k; ++j) a.z[i * N + j - 1] = 1.0;
        }
        return a;
    }
    Aff V(int i, int k) const {
        Aff a = zero();
        if (k == 0) a.p[2 * i + 1] = 1.0; else a.z[i * N + k - 1] = 1.0;
        return a;
    }
    Aff E(int e) const { Aff a = zero(); a.z[nl * N + e] = 1.0; return a; }
    Aff PAR(int t) const { Aff a = zero(); a.p[2 * nl + t] = 1.0; return a; }
    // entry (row, k) of the b-th (2, N+1) parameter block
    Aff PB(int b, int row, int k) const { return PAR(b * 2 * (N + 1) + row * (N + 1) + k); }
    Aff K(double c) const { Aff a = zero(); a.p[npv - 1] = c; return a; }
    // a cost term on the expression e: w e^2 (2-norm), or w |e| = w max(0, e) + w max(0, -e) as a pair of L1-penalised
    // rows (1-norm: min_1_norm of dmpcpwa's MpcMld, ||Q e||_1 with the diagonal Q)
    void residual(const Aff& e, double w) {
        if (!one_norm) { res.push_back(e); wres.push_back(w); return; }
        rows.push_back(e); wmax.push_back(w);
        rows.push_back((-1.0) * e); wmax.push_back(w);
    }
    void product(const Aff& a, const Aff& b) { la.push_back(a); lb.push_back(b); }
    void row(const Aff& e, double w) { rows.push_back(e); wmax.push_back(w); }
};

const double QXP = 1.0, QXV = 0.1, QU = 1.0, WSL = 1e4, DSAFE = 25.0;   // Params (common_controller_params.py:14-23)

// ||x - y - sigma(x)||^2_Qx with sigma(x) = [-t0 v - d0, 0]  (spacing_policy.py:13-37)
void track(Builder& B, const Aff& xp, const Aff& xv, const Aff& yp, const Aff& yv, double d0, double t0) {
    B.residual(xp - yp + t0 * xv + B.K(d0), QXP);
    B.residual(xv - yv, QXV);
}

int build_formulation(const hvp_mpc_desc& d, Builder*& out) {
    const int N = d.N, np1 = N + 1;
    const bool front = d.flags & HVP_FRONT, leader = d.flags & HVP_LEADER, trailer = d.flags & HVP_TRAILER;
    const bool real_ref = d.flags & HVP_REAL_VEHICLE_REF;
    const double d0 = d.d0, t0 = d.t0;
    Builder* Bp = nullptr;
    if (d.one_norm && d.kind != HVP_MPC_CENT && d.kind != HVP_MPC_LOCAL)
        return fail(-4, "mpc_create: the 1-norm cost is built for the centralized and the local formulation (kind %d asked)", d.kind);
    switch (d.kind) {
        case HVP_MPC_CENT: {          // mpcs/cent_mld.py:48-177
            const int n = d.n_local, L = d.leader_index;
            if (L < 0 || L >= n) return fail(-4, "mpc_create: leader_index %d out of range", L);
            if (real_ref && L != 0) return fail(-4, "mpc_create: real_vehicle_as_reference needs leader_index 0 (cent_mld.py:63-66)");
            Bp = new Builder(n, N, 0, 2 * np1);
            Builder& B = *Bp;
            B.one_norm = d.one_norm != 0;
            for (int k = 0; k <= N; ++k) {
                if (!real_ref) {      // :83-92
                    B.residual(B.P(L, k) - B.PB(0, 0, k), QXP);
                    B.residual(B.V(L, k) - B.PB(0, 1, k), QXV);
                } else {              // :93-105
                    track(B, B.P(0, k), B.V(0, k), B.PB(0, 0, k), B.PB(0, 1, k), d0, t0);
                }
            }
            for (int i = 1; i < n; ++i)   // :106-117
                for (int k = 0; k <= N; ++k) track(B, B.P(i, k), B.V(i, k), B.P(i - 1, k), B.V(i - 1, k), d0, t0);
            if (real_ref)             // :162-169
                for (int k = 0; k <= N; ++k) B.row(B.P(0, k) - B.PB(0, 0, k) + B.K(DSAFE), WSL);
            for (int i = 1; i < n; ++i)   // :170-177
                for (int k = 0; k <= N; ++k) B.row(B.P(i, k) - B.P(i - 1, k) + B.K(DSAFE), WSL);
        } break;
        case HVP_MPC_LOCAL: {         // fleet_decent_mld.py:61-208, fleet_seq_mld.py:63-219
            Bp = new Builder(1, N, 0, 3 * 2 * np1);
            Builder& B = *Bp;
            B.one_norm = d.one_norm != 0;
            for (int k = 0; k <= N; ++k) {
                if (!front && !leader) track(B, B.P(0, k), B.V(0, k), B.PB(0, 0, k), B.PB(0, 1, k), d0, t0);
                if (!trailer && !leader) track(B, B.PB(1, 0, k), B.PB(1, 1, k), B.P(0, k), B.V(0, k), d0, t0);
                if (leader) {
                    if (!real_ref) {
                        B.residual(B.P(0, k) - B.PB(2, 0, k), QXP);
                        B.residual(B.V(0, k) - B.PB(2, 1, k), QXV);
                    } else {
                        track(B, B.P(0, k), B.V(0, k), B.PB(2, 0, k), B.PB(2, 1, k), d0, t0);
                    }
                }
            }
            for (int k = 0; k <= N; ++k) {
                if (!front) B.row(B.P(0, k) - B.PB(0, 0, k) + B.K(DSAFE), WSL);
                if (!trailer) B.row(B.PB(1, 0, k) + B.K(DSAFE) - B.P(0, k), WSL);
                // fleet_seq_mld.py:211-219 (shares s_front with safe_front; a leader that is not the
                // front vehicle would couple the two rows through one slack -- not representable here)
                if (leader && real_ref && front) B.row(B.P(0, k) - B.PB(2, 0, k) + B.K(DSAFE), WSL);
            }
            if (leader && real_ref && !front)
                return fail(-4, "mpc_create: real_vehicle_as_reference needs the leader at the front");
        } break;
        case HVP_MPC_EVENT: {         // fleet_event_based.py:72-291
            const int nf = d.n_front, nb = d.n_behind;
            if (nf < 0 || nf > 2 || nb < 0 || nb > 2) return fail(-4, "mpc_create: n_front/n_behind must be 0..2");
            const int nl = (nf > 0) + 1 + (nb > 0);
            if (d.n_local != nl) return fail(-4, "mpc_create: event problem with n_front=%d n_behind=%d has %d local vehicles, got %d", nf, nb, nl, d.n_local);
            const int me = nf > 0 ? 1 : 0, f1 = me - 1, b1 = me + 1;
            Bp = new Builder(nl, N, 0, 3 * 2 * np1);
            Builder& B = *Bp;
            const int rl = d.leader_index;
            if (rl != HVP_NO_LEADER) {    // :146-165
                int who;
                if (rl == -1) who = b1; else if (rl == 0) who = me; else if (rl == 1) who = f1;
                else return fail(-4, "mpc_create: rel leader index must be -1, 0, or 1 (fleet_event_based.py:163-165)");
                if (who < 0 || who >= nl) return fail(-4, "mpc_create: rel_leader_index %d refers to a vehicle outside the local problem", rl);
                for (int k = 0; k <= N; ++k) {
                    B.residual(B.P(who, k) - B.PB(0, 0, k), QXP);
                    B.residual(B.V(who, k) - B.PB(0, 1, k), QXV);
                }
            }
            for (int k = 0; k <= N; ++k) {
                if (nf > 0) track(B, B.P(me, k), B.V(me, k), B.P(f1, k), B.V(f1, k), d0, t0);                    // :168-176
                if (nf > 1) track(B, B.P(f1, k), B.V(f1, k), B.PB(1, 0, k), B.PB(1, 1, k), d0, t0);              // :177-186
                if (nb > 0) track(B, B.P(b1, k), B.V(b1, k), B.P(me, k), B.V(me, k), d0, t0);                    // :188-199
                if (nb > 1) track(B, B.PB(2, 0, k), B.PB(2, 1, k), B.P(b1, k), B.V(b1, k), d0, t0);              // :200-212
            }
            for (int k = 0; k <= N; ++k) {   // :256-291
                if (nf > 0) B.row(B.P(me, k) - B.P(f1, k) + B.K(DSAFE), WSL);
                if (nf > 1) B.row(B.P(f1, k) - B.PB(1, 0, k) + B.K(DSAFE), WSL);
                if (nb > 0) B.row(B.P(b1, k) - B.P(me, k) + B.K(DSAFE), WSL);
                if (nb > 1) B.row(B.PB(2, 0, k) - B.P(b1, k) + B.K(DSAFE), WSL);
            }
        } break;
        case HVP_MPC_ADMM: {          // fleet_naive_admm.py:63-237
            const int ne = (front ? 0 : 2 * np1) + (trailer ? 0 : 2 * np1);
            Bp = new Builder(1, N, ne, 5 * 2 * np1);
            Builder& B = *Bp;
            const int ef = 0, eb = front ? 0 : 2 * np1;     // extras: x_front (2,N+1) then x_back (2,N+1)
            const double rho = d.rho;
            if (!(rho > 0)) return fail(-4, "mpc_create: ADMM needs rho > 0");
            for (int k = 0; k <= N; ++k) {
                if (!front && !leader) track(B, B.P(0, k), B.V(0, k), B.E(ef + k), B.E(ef + np1 + k), d0, t0);        // :127-138
                if (!trailer && !leader) track(B, B.E(eb + k), B.E(eb + np1 + k), B.P(0, k), B.V(0, k), d0, t0);      // :139-150
                if (leader) {                                                                                      // :151-157
                    B.residual(B.P(0, k) - B.PB(0, 0, k), QXP);
                    B.residual(B.V(0, k) - B.PB(0, 1, k), QXV);
                }
                for (int row = 0; row < 2; ++row) {
                    if (!front) {     // :174-186: y'(x_front - z) + rho/2 ||x_front - z||^2
                        const Aff dlt = B.E(ef + row * np1 + k) - B.PB(2, row, k);
                        B.product(B.PB(1, row, k), dlt);
                        B.residual(dlt, 0.5 * rho);
                    }
                    if (!trailer) {   // :187-199
                        const Aff dlt = B.E(eb + row * np1 + k) - B.PB(4, row, k);
                        B.product(B.PB(3, row, k), dlt);
                        B.residual(dlt, 0.5 * rho);
                    }
                }
            }
            for (int k = 0; k <= N; ++k) {   // :219-237
                if (!front) B.row(B.P(0, k) - B.E(ef + k) + B.K(DSAFE), WSL);
                if (!trailer) B.row(B.E(eb + k) + B.K(DSAFE) - B.P(0, k), WSL);
            }
        } break;
        case HVP_MPC_GADMM: {         // fleet_g_admm.py:55-158 (+ dmpcrl MpcAdmm augmented state / consensus terms)
            const int nf = d.n_front, nb = d.n_behind, nc = nf + nb, na = nc + 1;
            if (nf < 0 || nb < 0 || nc > 4) return fail(-4, "mpc_create: GADMM copies out of range");
            if (!leader && nf < 1) return fail(-4, "mpc_create: a GADMM follower tracks its first copy (fleet_g_admm.py:136-158): n_front >= 1");
            Bp = new Builder(1, N, nc * 2 * np1, (1 + 2 * na) * 2 * np1);
            Builder& B = *Bp;
            const double rho = d.rho;
            if (!(rho > 0)) return fail(-4, "mpc_create: GADMM needs rho > 0");
            // copy c (0..nc-1) entry (row,k); augmented index of copy c: c < nf ? c : c + 1; own state: nf
            auto CP = [&](int c, int row, int k) { return B.E(c * 2 * np1 + row * np1 + k); };
            auto YZ = [&](int which, int a, int row, int k) {       // which: 0 = y, 1 = z; a: augmented index
                return B.PAR(2 * np1 + which * na * 2 * np1 + (2 * a + row) * np1 + k);
            };
            for (int k = 0; k <= N; ++k) {
                if (leader) {         // :112-135
                    B.residual(B.P(0, k) - B.PB(0, 0, k), QXP);
                    B.residual(B.V(0, k) - B.PB(0, 1, k), QXV);
                } else {              // :136-158: follows the FIRST copy
                    track(B, B.P(0, k), B.V(0, k), CP(0, 0, k), CP(0, 1, k), d0, t0);
                    B.row(B.P(0, k) - CP(0, 0, k) + B.K(DSAFE), WSL);        // :98-109
                }
                for (int a = 0; a < na; ++a)
                    for (int row = 0; row < 2; ++row) {
                        Aff xa = (a == nf) ? (row == 0 ? B.P(0, k) : B.V(0, k)) : CP(a < nf ? a : a - 1, row, k);
                        const Aff dlt = xa - YZ(1, a, row, k);
                        B.product(YZ(0, a, row, k), dlt);
                        B.residual(dlt, 0.5 * rho);
                    }
            }
        } break;
        default:
            return fail(-4, "mpc_create: unknown kind %d", d.kind);
    }
    out = Bp;
    return 0;
}

void fill_model(PmModel& M, int model) {
    VehicleModel V;
    const double beta = (3 * V.c_fric * V.v_max * V.v_max) / 16;   // models.py:276-282
    const double alpha = V.v_max / 2;
    const double c1 = beta / alpha;
    const double c2 = (V.c_fric * V.v_max * V.v_max - beta) / (V.v_max - alpha);
    const double dfr = beta - alpha * ((V.c_fric * V.v_max * V.v_max - beta) / (V.v_max - alpha));
    M.mug = V.mu * V.grav;
    if (model == HVP_MODEL_PWA_GEAR) {          // models.py:397-492
        double lim[5];
        V.gear_limits(lim);
        const double e[8] = {-HUGE_VAL, lim[0], lim[1], lim[2], alpha, lim[3], lim[4], HUGE_VAL};
        const int g[7] = {0, 1, 2, 3, 3, 4, 5};
        M.R = 7;
        for (int r = 0; r < 7; ++r) {
            M.cf[r] = r < 4 ? c1 : c2; M.dd[r] = r < 4 ? 0.0 : dfr; M.bg[r] = V.bgear[g[r]];
            M.lo[r] = e[r]; M.hi[r] = e[r + 1]; M.gear[r] = g[r] + 1;
        }
    } else {                                    // pwa_friction (models.py:288-332) x gears (mpc_gear.py:30-114)
        M.R = 12;
        for (int f = 0; f < 2; ++f)
            for (int j = 0; j < 6; ++j) {
                const int r = f * 6 + j;
                M.cf[r] = f == 0 ? c1 : c2; M.dd[r] = f == 0 ? 0.0 : dfr; M.bg[r] = V.bgear[j];
                M.lo[r] = fmax(f == 0 ? -HUGE_VAL : alpha, V.vl[j]);
                M.hi[r] = fmin(f == 0 ? alpha : HUGE_VAL, V.vh[j]);
                M.gear[r] = j + 1;
            }
    }
    for (int r = M.R; r < PM_MAXMODES; ++r) { M.cf[r] = M.bg[r] = M.dd[r] = 0; M.lo[r] = 1; M.hi[r] = 0; M.gear[r] = 0; }
}

// inverse of a symmetric positive definite matrix (Cholesky), row-major n x n; returns false if not PD
bool spd_inverse(int n, const std::vector<double>& A, std::vector<double>& inv) {
    std::vector<double> L(A);
    for (int j = 0; j < n; ++j) {
        double dg = L[j * n + j];
        for (int k = 0; k < j; ++k) dg -= L[j * n + k] * L[j * n + k];
        if (!(dg > 0)) return false;
        dg = sqrt(dg);
        L[j * n + j] = dg;
        for (int i = j + 1; i < n; ++i) {
            double s = L[i * n + j];
            for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = s / dg;
        }
    }
    inv.assign((size_t)n * n, 0.0);
    std::vector<double> y(n);
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < n; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = 0; k < i; ++k) s -= L[i * n + k] * y[k];
            y[i] = s / L[i * n + i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = y[i];
            for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * y[k];
            y[i] = s / L[i * n + i];
        }
        for (int i = 0; i < n; ++i) inv[(size_t)i * n + c] = y[i];
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) {
            const double s = 0.5 * (inv[(size_t)i * n + j] + inv[(size_t)j * n + i]);
            inv[(size_t)i * n + j] = inv[(size_t)j * n + i] = s;
        }
    return true;
}

}  // namespace

static int upload(hvp_mpc* m, const std::vector<double>& h, const double** out) {
    void* p = nullptr;
    const size_t bytes = (h.empty() ? 1 : h.size()) * sizeof(double);
    CUDA_TRY(cudaMalloc(&p, bytes));
    m->dev.push_back(p);
    if (!h.empty()) CUDA_TRY(cudaMemcpy(p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    *out = (const double*)p;
    return 0;
}

extern "C" int hvp_mpc_destroy(hvp_mpc* m) {
    if (!m) return 0;
    HvpRelaxedCapture relaxed__;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    for (void* p : m->dev) cudaFree(p);
    if (m->x0buf) cudaFree(m->x0buf);
    if (m->ybuf) cudaFree(m->ybuf);
    if (m->scratch_mem) cudaFree(m->scratch_mem);
    if (m->shard_mem) cudaFree(m->shard_mem);
    delete m;
    return 0;
}

extern "C" int hvp_mpc_create(hvp_ctx* c, const hvp_mpc_desc* d, hvp_mpc** out) {
    if (!c || !d || !out) return fail(-1, "mpc_create: NULL argument");
    if (d->N < 2 || d->N > 16) return fail(-4, "mpc_create: N=%d out of range [2,16]", d->N);
    if (d->n_local < 1 || d->n_local > 16) return fail(-4, "mpc_create: n_local=%d out of range [1,16]", d->n_local);
    if (d->model != HVP_MODEL_PWA_GEAR && d->model != HVP_MODEL_FRICTION_GEAR)
        return fail(-4, "mpc_create: unknown model %d", d->model);
    if (d->kind != HVP_MPC_CENT && d->kind != HVP_MPC_EVENT && d->n_local != 1)
        return fail(-4, "mpc_create: kind %d has exactly one local vehicle", d->kind);
    if (d->max_nodes < 0) return fail(-4, "mpc_create: negative max_nodes");
    if (!(d->mip_gap >= 0.0) || !(d->mip_gap < 1.0)) return fail(-4, "mpc_create: mip_gap must be in [0, 1)");
    if (!(d->time_limit_ms >= 0.0)) return fail(-4, "mpc_create: negative time_limit_ms");
    Builder* Bp = nullptr;
    int rc = build_formulation(*d, Bp);
    if (rc) { delete Bp; return rc; }
    Builder& B = *Bp;
    hvp_mpc* m = new hvp_mpc();
    m->ctx = c; m->desc = *d;
    PmDev& S = m->S;
    memset(&S, 0, sizeof S);
    S.nl = B.nl; S.N = B.N; S.ne = B.ne; S.nv = B.nv; S.npar = B.npar; S.npv = B.npv;
    S.depth = B.nl * B.N; S.max_nodes = d->max_nodes;
    S.mip_gap = d->mip_gap; S.time_limit_ns = (long long)(d->time_limit_ms * 1e6);
    fill_model(S.M, d->model);
    VehicleModel V;
    S.qu = QU; S.w = WSL; S.vmin = V.v_min; S.vmax = V.v_max; S.pmin = V.p_min; S.pmax = V.p_max;
    S.umin = V.u_min; S.umax = V.u_max; S.a_acc = 2.5; S.a_dec = -2.0; S.tight = d->tight;
    const int nv = B.nv, npv = B.npv;
    // split rows into generic (non-zero normal) and constant ones
    std::vector<int> gen, cst;
    for (size_t r = 0; r < B.rows.size(); ++r) (B.rows[r].has_z() ? gen : cst).push_back((int)r);
    S.nres = (int)B.res.size(); S.nlin = (int)B.la.size(); S.ng = (int)gen.size(); S.n0 = (int)cst.size();
    if (nv > 64 || S.ng > 4000) { delete Bp; delete m; return fail(-4, "mpc_create: problem too large (nv=%d, rows=%d; max 64 variables)", nv, S.ng); }
    std::vector<double> H0((size_t)nv * nv, 0.0), H0inv, Cres((size_t)S.nres * npv), wres(S.nres), RW2((size_t)nv * S.nres);
    for (int r = 0; r < S.nres; ++r) {
        const Aff& e = B.res[r];
        const double w = B.wres[r];
        wres[r] = w;
        for (int t = 0; t < npv; ++t) Cres[(size_t)r * npv + t] = e.p[t];
        for (int i = 0; i < nv; ++i) {
            RW2[(size_t)i * S.nres + r] = 2.0 * w * e.z[i];
            if (e.z[i] == 0.0) continue;
            for (int j = 0; j < nv; ++j) H0[(size_t)i * nv + j] += 2.0 * w * e.z[i] * e.z[j];
        }
    }
    S.sibling = getenv("HVP_MPC_SIBLING") ? atoi(getenv("HVP_MPC_SIBLING")) : 1;
    S.one_norm = B.one_norm ? 1 : 0;
    // rho: small enough that a round reaches the solution set in one or two steps (step ~ cost slope / rho), large
    // enough that saturating a w = 1e4 slack row does not send the iterate through 1e8 (round-off 1e-8: measured, 2.5 %
    // of the C1-shaped problems did not settle at rho = 1e-4, 0.1 % at 1e-3 .. 1e-2 before the objective criterion)
    S.rho_px = getenv("HVP_RHO_PX") ? atof(getenv("HVP_RHO_PX")) : 1e-3;
    S.ppa_stall = getenv("HVP_PPA_STALL") ? atof(getenv("HVP_PPA_STALL")) : 1e-12;
    if (B.one_norm)                    // no quadratic cost term: the proximal term of the node LPs (pm_types.h)
        for (int j = 0; j < nv; ++j) H0[(size_t)j * nv + j] = S.rho_px;
    if (!spd_inverse(nv, H0, H0inv)) {
        delete Bp; delete m;
        return fail(-5, "mpc_create: the tracking Hessian of this formulation is not positive definite");
    }
    std::vector<double> La((size_t)S.nlin * npv), Lz((size_t)S.nlin * nv), Lp((size_t)S.nlin * npv);
    for (int l = 0; l < S.nlin; ++l) {
        if (B.la[l].has_z()) { delete Bp; delete m; return fail(-5, "mpc_create: internal: bilinear cost term"); }
        for (int t = 0; t < npv; ++t) { La[(size_t)l * npv + t] = B.la[l].p[t]; Lp[(size_t)l * npv + t] = B.lb[l].p[t]; }
        for (int j = 0; j < nv; ++j) Lz[(size_t)l * nv + j] = B.lb[l].z[j];
    }
    std::vector<double> AT((size_t)nv * S.ng), BR((size_t)S.ng * npv), wmax(S.ng), B0((size_t)S.n0 * npv), w0(S.n0);
    for (int g = 0; g < S.ng; ++g) {
        const Aff& e = B.rows[gen[g]];
        for (int j = 0; j < nv; ++j) AT[(size_t)j * S.ng + g] = e.z[j];
        for (int t = 0; t < npv; ++t) BR[(size_t)g * npv + t] = -e.p[t];
        wmax[g] = B.wmax[gen[g]];
    }
    for (int g = 0; g < S.n0; ++g) {
        const Aff& e = B.rows[cst[g]];
        for (int t = 0; t < npv; ++t) B0[(size_t)g * npv + t] = e.p[t];
        w0[g] = B.wmax[cst[g]];
    }
    // gradient map of the FULL variable vector: g = G . pvec, G = RW2.Cres + Lz'.La
    std::vector<double> G((size_t)nv * npv, 0.0);
    for (int j = 0; j < nv; ++j)
        for (int t = 0; t < npv; ++t) {
            double s = 0.0;
            for (int r = 0; r < S.nres; ++r) s += RW2[(size_t)j * S.nres + r] * Cres[(size_t)r * npv + t];
            for (int l = 0; l < S.nlin; ++l) s += Lz[(size_t)l * nv + j] * La[(size_t)l * npv + t];
            G[(size_t)j * npv + t] = s;
        }
    // ---- elimination of the free copies that appear in no row (pm_types.h) ----
    std::vector<int> emap(B.ne > 0 ? B.ne : 1, 0);
    std::vector<double> Rz, GE, HE;
    int nk = nv;
    {
        static const int elim_on = getenv("HVP_MPC_ELIM") ? atoi(getenv("HVP_MPC_ELIM")) : 1;
        std::vector<int> K, E;
        for (int j = 0; j < nv; ++j) {
            bool in_row = j < B.nl * B.N;
            for (int g = 0; g < S.ng && !in_row; ++g) in_row = AT[(size_t)j * S.ng + g] != 0.0;
            (in_row || !elim_on ? K : E).push_back(j);
        }
        const int nel = (int)E.size();
        nk = (int)K.size();
        int kept = 0, gone = 0;
        for (int e = 0; e < B.ne; ++e) {
            const int j = B.nl * B.N + e;
            bool is_e = false;
            for (int q : E) is_e = is_e || q == j;
            emap[e] = is_e ? -1 - gone++ : kept++;
        }
        if (nel > 0) {
            std::vector<double> C((size_t)nel * nel), Cinv, Bm((size_t)nk * nel);
            for (int a = 0; a < nel; ++a)
                for (int b = 0; b < nel; ++b) C[(size_t)a * nel + b] = H0[(size_t)E[a] * nv + E[b]];
            if (!spd_inverse(nel, C, Cinv)) { delete Bp; delete m; return fail(-5, "mpc_create: eliminated block is not positive definite"); }
            for (int a = 0; a < nk; ++a)
                for (int b = 0; b < nel; ++b) Bm[(size_t)a * nel + b] = H0[(size_t)K[a] * nv + E[b]];
            std::vector<double> BC((size_t)nk * nel, 0.0);                 // B C^-1
            for (int a = 0; a < nk; ++a)
                for (int b = 0; b < nel; ++b) {
                    double s2 = 0.0;
                    for (int c2 = 0; c2 < nel; ++c2) s2 += Bm[(size_t)a * nel + c2] * Cinv[(size_t)c2 * nel + b];
                    BC[(size_t)a * nel + b] = s2;
                }
            std::vector<double> H0r((size_t)nk * nk), Gr((size_t)nk * npv), ATr((size_t)nk * S.ng);
            for (int a = 0; a < nk; ++a) {
                for (int b = 0; b < nk; ++b) {
                    double s2 = H0[(size_t)K[a] * nv + K[b]];
                    for (int c2 = 0; c2 < nel; ++c2) s2 -= BC[(size_t)a * nel + c2] * Bm[(size_t)b * nel + c2];
                    H0r[(size_t)a * nk + b] = s2;
                }
                for (int t = 0; t < npv; ++t) {
                    double s2 = G[(size_t)K[a] * npv + t];
                    for (int c2 = 0; c2 < nel; ++c2) s2 -= BC[(size_t)a * nel + c2] * G[(size_t)E[c2] * npv + t];
                    Gr[(size_t)a * npv + t] = s2;
                }
                for (int g = 0; g < S.ng; ++g) ATr[(size_t)a * S.ng + g] = AT[(size_t)K[a] * S.ng + g];
            }
            for (int a = 0; a < nk; ++a)                                   // symmetrise the round-off
                for (int b = 0; b < a; ++b) {
                    const double s2 = 0.5 * (H0r[(size_t)a * nk + b] + H0r[(size_t)b * nk + a]);
                    H0r[(size_t)a * nk + b] = H0r[(size_t)b * nk + a] = s2;
                }
            GE.assign((size_t)nel * npv, 0.0); HE.assign((size_t)nel * npv, 0.0); Rz.assign((size_t)nel * nk, 0.0);
            for (int a = 0; a < nel; ++a) {
                for (int t = 0; t < npv; ++t) {
                    GE[(size_t)a * npv + t] = G[(size_t)E[a] * npv + t];
                    double s2 = 0.0;
                    for (int c2 = 0; c2 < nel; ++c2) s2 -= Cinv[(size_t)a * nel + c2] * G[(size_t)E[c2] * npv + t];
                    HE[(size_t)a * npv + t] = s2;
                }
                for (int b = 0; b < nk; ++b) Rz[(size_t)a * nk + b] = -BC[(size_t)b * nel + a];   // -(C^-1 B')[a][b], C^-1 symmetric
            }
            // the kernel works on the kept variables; the residual maps of the fallback path (no precompute) do not
            // carry the elimination, so they are cut down only to keep the indexing consistent
            std::vector<double> RW2r((size_t)nk * S.nres), Lzr((size_t)S.nlin * nk);
            for (int a = 0; a < nk; ++a)
                for (int r = 0; r < S.nres; ++r) RW2r[(size_t)a * S.nres + r] = RW2[(size_t)K[a] * S.nres + r];
            for (int l = 0; l < S.nlin; ++l)
                for (int a = 0; a < nk; ++a) Lzr[(size_t)l * nk + a] = Lz[(size_t)l * nv + K[a]];
            H0.swap(H0r); G.swap(Gr); AT.swap(ATr); RW2.swap(RW2r); Lz.swap(Lzr);
            if (!spd_inverse(nk, H0, H0inv)) { delete Bp; delete m; return fail(-5, "mpc_create: reduced Hessian is not positive definite"); }
            S.nv = nk;
        }
        S.nel = nel;
    }
    // stacked matrix of the tensor-core precompute (pm_types.h)
    {
        const int nres = S.nres, nlin = S.nlin, nel = S.nel;
        S.kw = (npv + 3) / 4 * 4;
        S.mw = (nk + S.ng + S.n0 + nres + 2 * nlin + 2 * nel + 7) / 8 * 8;
        std::vector<double> Wm((size_t)S.mw * S.kw, 0.0);
        for (int j = 0; j < nk; ++j)
            for (int t = 0; t < npv; ++t) Wm[(size_t)j * S.kw + t] = G[(size_t)j * npv + t];
        int row = nk;
        for (int g = 0; g < S.ng; ++g, ++row)
            for (int t = 0; t < npv; ++t) Wm[(size_t)row * S.kw + t] = BR[(size_t)g * npv + t];
        for (int g = 0; g < S.n0; ++g, ++row)
            for (int t = 0; t < npv; ++t) Wm[(size_t)row * S.kw + t] = B0[(size_t)g * npv + t];
        // the constant term is summed from the residuals themselves (c0 = sum w_r c_r^2 + sum a_l p_l): the
        // residuals are small numbers, whereas the quadratic form pvec' CC pvec cancels terms of size p^2 ~ 1e7
        for (int r = 0; r < nres; ++r, ++row)
            for (int t = 0; t < npv; ++t) Wm[(size_t)row * S.kw + t] = Cres[(size_t)r * npv + t];
        for (int l = 0; l < nlin; ++l, ++row)
            for (int t = 0; t < npv; ++t) Wm[(size_t)row * S.kw + t] = La[(size_t)l * npv + t];
        for (int l = 0; l < nlin; ++l, ++row)
            for (int t = 0; t < npv; ++t) Wm[(size_t)row * S.kw + t] = Lp[(size_t)l * npv + t];
        S.o_yel = row;
        for (int e = 0; e < nel; ++e, ++row)
            for (int t = 0; t < npv; ++t) Wm[(size_t)row * S.kw + t] = GE[(size_t)e * npv + t];
        for (int e = 0; e < nel; ++e, ++row)
            for (int t = 0; t < npv; ++t) Wm[(size_t)row * S.kw + t] = HE[(size_t)e * npv + t];
        rc = upload(m, Wm, &S.W);
        if (rc) { delete Bp; hvp_mpc_destroy(m); return rc; }
    }
    delete Bp;
    pm_layout(S);
    if (S.smem_bytes > 220 * 1024) { delete m; return fail(-4, "mpc_create: problem needs %d bytes of shared memory per warp", S.smem_bytes); }
    cudaSetDevice(c->device);
    rc = upload(m, H0, &S.H0); if (!rc) rc = upload(m, H0inv, &S.H0inv);
    if (!rc) rc = upload(m, Cres, &S.Cres); if (!rc) rc = upload(m, wres, &S.wres); if (!rc) rc = upload(m, RW2, &S.RW2);
    if (!rc) rc = upload(m, La, &S.La); if (!rc) rc = upload(m, Lz, &S.Lz); if (!rc) rc = upload(m, Lp, &S.Lp);
    if (!rc) rc = upload(m, AT, &S.AT); if (!rc) rc = upload(m, BR, &S.BR); if (!rc) rc = upload(m, wmax, &S.wmax);
    if (!rc) rc = upload(m, B0, &S.B0); if (!rc) rc = upload(m, w0, &S.w0);
    if (!rc) rc = upload(m, Rz, &S.Rz);
    if (!rc) {
        void* pe = nullptr;
        if (cudaMalloc(&pe, emap.size() * sizeof(int)) != cudaSuccess) rc = fail(-100, "mpc_create: cudaMalloc failed");
        else { m->dev.push_back(pe); cudaMemcpy(pe, emap.data(), emap.size() * sizeof(int), cudaMemcpyHostToDevice); S.emap = (const int*)pe; }
    }
    if (!rc) { const double* cn = nullptr; rc = upload(m, std::vector<double>(2, 0.0), &cn); m->counter = (unsigned long long*)cn; }
    if (rc) { hvp_mpc_destroy(m); return rc; }
    // tree splitting of heavy problems (PmSplit): prefix depth of the split; the scratch is sized per batch
    memset(&m->scratch, 0, sizeof m->scratch);
    memset(&m->shard, 0, sizeof m->shard);
    {
        const char* env = getenv("HVP_MPC_SPLIT");
        // prefix levels are enumerated, never solved, so nothing prunes inside them: deep prefixes multiply the node count
        // (centralized n = 10, N = 6: 280 000 nodes per MIQP at depth 30, 2 000 at depth 16; 1-norm n = 3, N = 5: 370 at depth
        // 9, 223 at depth 5) and, with waiting workers adopting sub-trees, are no longer needed for balance
        int D = S.nl + 2 > 5 ? S.nl + 2 : 5;
        if (D > S.depth - 1) D = S.depth - 1;
        if (env && atoi(env) == 0) D = 0;
        else if (env && atoi(env) > 1) D = atoi(env) < S.depth - 1 ? atoi(env) : S.depth - 1;   // explicit prefix depth
        m->split_D = D;
    }
    *out = m;
    return 0;
}

extern "C" int hvp_mpc_info(const hvp_mpc* m, int32_t* info) {
    if (!m || !info) return fail(-1, "mpc_info: NULL argument");
    const PmDev& S = m->S;
    info[0] = S.nv; info[1] = S.ne; info[2] = S.npar; info[3] = S.M.R; info[4] = S.nl; info[5] = S.N;
    info[6] = S.ng; info[7] = S.smem_bytes;
    return 0;
}

extern "C" int hvp_mpc_mode_table(const hvp_mpc* m, double* lo, double* hi, int32_t* gear) {
    if (!m) return fail(-1, "mpc_mode_table: NULL argument");
    for (int r = 0; r < m->S.M.R; ++r) {
        if (lo) lo[r] = m->S.M.lo[r];
        if (hi) hi[r] = m->S.M.hi[r];
        if (gear) gear[r] = m->S.M.gear[r];
    }
    return 0;
}

extern "C" int hvp_mpc_solve_dev(hvp_mpc* m, int64_t batch, const double* x0, const double* mass,
                                 const double* params, const int32_t* fixed_modes, double* u, double* x,
                                 double* extra, int32_t* modes, double* obj, int32_t* status, int32_t* nodes,
                                 int32_t* qp_iters, void* stream) {
    if (!m) return fail(-1, "hvp_mpc_solve_dev: NULL handle");
    std::unique_lock<std::mutex> lk__(m->mu, std::defer_lock);
    if (!hvp_mpc_locked_by_this_thread) lk__.lock();
    HvpReentry re__;
    if (!m) return fail(-1, "mpc_solve: handle is NULL");
    if (batch < 0) return fail(-4, "mpc_solve: negative batch");
    if (batch == 0) return 0;
    if (!x0 || !mass || !params || !u || !x || !modes || !obj || !status || !nodes)
        return fail(-1, "mpc_solve: NULL array argument");
    hvp_ctx* c = m->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const size_t need = (size_t)batch * m->S.mw * sizeof(double);
    if (need > m->ycap) {
        if (m->ybuf) { CUDA_TRY(cudaStreamSynchronize(st)); CUDA_TRY(cudaFree(m->ybuf)); m->ybuf = nullptr; m->ycap = 0; }
        CUDA_TRY(cudaMalloc(&m->ybuf, need + need / 4));
        m->ycap = need + need / 4;
    }
    if (m->split_D >= 1 && !fixed_modes && m->S.max_nodes == 0) {
        // cap heavy problems x M warps each; sized from the batch (grow-only)
        static const int envM = getenv("HVP_MPC_SPLIT_M") ? atoi(getenv("HVP_MPC_SPLIT_M")) : 64;
        static const int envB = getenv("HVP_MPC_BUDGET") ? atoi(getenv("HVP_MPC_BUDGET")) : 64;
        // Room for EVERY problem of the batch on the flagged list, as far as 512 MB of result slots go: a heavy tree
        // that finds the list full is finished by the single worker of the budgeted pass -- r02q launch list of the
        // N = 10 time-headway group: after the leader's speed change more than batch / 8 trees are heavy, and the
        // budgeted pass went from 8 ms to 60-106 ms (one 6 880-node tree on one 16-lane group).
        size_t cap = (size_t)batch;
        {
            const PmDev& S0 = m->S;
            const size_t nu0 = (size_t)S0.nl * S0.N, nx0 = (size_t)S0.nl * 2 * (S0.N + 1), ne0 = S0.ne > 0 ? (size_t)S0.ne : 1;
            const size_t slot = (nu0 + nx0 + ne0 + 1) * 8 + (nu0 + 3) * 4;
            size_t cap_max = ((size_t)512 << 20) / (slot * (size_t)envM);
            if (cap_max > 32768) cap_max = 32768;
            if (cap > cap_max) cap = cap_max;
        }
        if (cap < 256) cap = 256;
        if (cap > m->scratch_cap || m->scratch.sp.M != envM) {
            if (m->scratch_mem) { CUDA_TRY(cudaStreamSynchronize(st)); CUDA_TRY(cudaFree(m->scratch_mem)); m->scratch_mem = nullptr; }
            const PmDev& S = m->S;
            // result slots: cap * M work items + the pool of the adopted sub-trees (PmSplit::ad)
            const size_t items = cap * (size_t)envM + PM_POOL, nu = (size_t)S.nl * S.N, nx = (size_t)S.nl * 2 * (S.N + 1);
            const size_t ne = S.ne > 0 ? (size_t)S.ne : 1;
            const size_t bytes = items * ((nu + nx + ne + 1) * 8 + (nu + 3) * 4) + cap * (4 + 8) + 64 + 12 * 256 +
                                 pm_adopt_bytes(S.depth);
            void* p = nullptr;
            CUDA_TRY(cudaMalloc(&p, bytes));
            m->scratch_mem = p;
            char* q = (char*)p;
            auto take = [&](size_t b) { char* r = q; q += (b + 255) & ~(size_t)255; return r; };
            PmScratch& sc = m->scratch;
            sc.u = (double*)take(items * nu * 8); sc.x = (double*)take(items * nx * 8); sc.extra = (double*)take(items * ne * 8);
            sc.obj = (double*)take(items * 8); sc.modes = (int32_t*)take(items * nu * 4); sc.status = (int32_t*)take(items * 4);
            sc.nodes = (int32_t*)take(items * 4); sc.iters = (int32_t*)take(items * 4);
            sc.sp.inc_shared = (unsigned long long*)take(cap * 8); sc.sp.flagged = (int*)take(cap * 4);
            sc.sp.nflag = (int*)take(4);
            pm_adopt_take(sc.sp, take, S.depth);
            sc.sp.M = envM; sc.sp.D = m->split_D;
            m->scratch_cap = cap;
        }
        m->scratch.sp.cap = (int)cap;
        m->scratch.sp.budget = envB;
    }
    CUDA_TRY(hvp_mark(c, c->ev0, st, false));
    CUDA_TRY(launch_pm_precompute(m->S, batch, x0, params, m->ybuf, st));
    CUDA_TRY(launch_pm_miqp(m->S, batch, x0, mass, params, fixed_modes, m->ybuf, u, x, extra, modes, obj, status,
                            nodes, qp_iters, m->counter, (m->split_D >= 1 && m->scratch_mem) ? &m->scratch : nullptr, st));
    CUDA_TRY(hvp_mark(c, c->ev1, st, true));
    c->launches += 2;
    return 0;
}

extern "C" int hvp_mpc_solve_shard_dev(hvp_mpc* m, int64_t batch, const double* x0, const double* mass,
                                       const double* params, int32_t rank, int32_t world, int32_t groups,
                                       int32_t prefix_depth, int32_t node_budget, const double* incumbent, double* u,
                                       double* x, double* extra, int32_t* modes, double* obj, int32_t* status,
                                       int32_t* nodes, int32_t* qp_iters, void* stream) {
    if (!m) return fail(-1, "hvp_mpc_solve_shard_dev: NULL handle");
    std::unique_lock<std::mutex> lk__(m->mu, std::defer_lock);
    if (!hvp_mpc_locked_by_this_thread) lk__.lock();
    HvpReentry re__;
    if (!m) return fail(-1, "mpc_solve_shard: handle is NULL");
    if (batch < 0) return fail(-4, "mpc_solve_shard: negative batch");
    if (world < 1 || rank < 0 || rank >= world) return fail(-4, "mpc_solve_shard: rank %d outside world %d", rank, world);
    if (groups < 1 || groups > 4096) return fail(-4, "mpc_solve_shard: groups must be 1..4096");
    if (node_budget < 0) return fail(-4, "mpc_solve_shard: negative node budget");
    if (batch == 0) return 0;
    if (!x0 || !mass || !params || !u || !x || !modes || !obj || !status || !nodes)
        return fail(-1, "mpc_solve_shard: NULL array argument");
    const PmDev& S = m->S;
    if (S.max_nodes != 0) return fail(-4, "mpc_solve_shard: the handle was created with max_nodes");
    int D = prefix_depth > 0 ? prefix_depth : m->split_D;
    if (D > S.depth - 1) D = S.depth - 1;
    if (D < 1) return fail(-4, "mpc_solve_shard: the formulation has no tree to split (depth %d)", S.depth);
    if ((size_t)batch * (size_t)groups > ((size_t)1 << 24)) return fail(-4, "mpc_solve_shard: batch x groups too large");
    hvp_ctx* c = m->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const size_t need = (size_t)batch * S.mw * sizeof(double);
    if (need > m->ycap) {
        if (m->ybuf) { CUDA_TRY(cudaStreamSynchronize(st)); CUDA_TRY(cudaFree(m->ybuf)); m->ybuf = nullptr; m->ycap = 0; }
        CUDA_TRY(cudaMalloc(&m->ybuf, need + need / 4));
        m->ycap = need + need / 4;
    }
    const size_t items = (size_t)batch * (size_t)groups + PM_POOL;      // work items + the pool of the adopted sub-trees
    if (items > m->shard_items || (size_t)batch > (size_t)m->shard.sp.cap) {
        if (m->shard_mem) { CUDA_TRY(cudaStreamSynchronize(st)); CUDA_TRY(cudaFree(m->shard_mem)); m->shard_mem = nullptr; }
        const size_t nu = (size_t)S.nl * S.N, nx = (size_t)S.nl * 2 * (S.N + 1), ne = S.ne > 0 ? (size_t)S.ne : 1;
        const size_t bytes = items * ((nu + nx + ne + 1) * 8 + (nu + 3) * 4) + (size_t)batch * (4 + 8) + 64 + 12 * 256 +
                             pm_adopt_bytes(S.depth);
        void* p = nullptr;
        CUDA_TRY(cudaMalloc(&p, bytes));
        m->shard_mem = p;
        char* q = (char*)p;
        auto take = [&](size_t b) { char* r = q; q += (b + 255) & ~(size_t)255; return r; };
        PmScratch& sc = m->shard;
        sc.u = (double*)take(items * nu * 8); sc.x = (double*)take(items * nx * 8); sc.extra = (double*)take(items * ne * 8);
        sc.obj = (double*)take(items * 8); sc.modes = (int32_t*)take(items * nu * 4); sc.status = (int32_t*)take(items * 4);
        sc.nodes = (int32_t*)take(items * 4); sc.iters = (int32_t*)take(items * 4);
        sc.sp.inc_shared = (unsigned long long*)take((size_t)batch * 8); sc.sp.flagged = (int*)take((size_t)batch * 4);
        sc.sp.nflag = (int*)take(4);
        pm_adopt_take(sc.sp, take, S.depth);
        sc.sp.cap = (int)batch;
        m->shard_items = items;
    }
    PmScratch sc = m->shard;
    sc.sp.M = groups; sc.sp.D = D; sc.sp.cap = (int)batch; sc.sp.budget = node_budget;
    sc.sp.rank = rank; sc.sp.world = world;
    CUDA_TRY(hvp_mark(c, c->ev0, st, false));
    CUDA_TRY(launch_pm_precompute(S, batch, x0, params, m->ybuf, st));
    CUDA_TRY(launch_pm_shard(S, batch, x0, mass, params, m->ybuf, incumbent, u, x, extra, modes, obj, status, nodes,
                             qp_iters, m->counter, &sc, st));
    CUDA_TRY(hvp_mark(c, c->ev1, st, true));
    c->launches += 4;
    return 0;
}

static inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

extern "C" int hvp_mpc_solve_host(hvp_mpc* m, int64_t batch, const double* x0, const double* mass,
                                  const double* params, const int32_t* fixed_modes, double* u, double* x,
                                  double* extra, int32_t* modes, double* obj, int32_t* status, int32_t* nodes,
                                  int32_t* qp_iters) {
    if (!m) return fail(-1, "hvp_mpc_solve_host: NULL handle");
    std::unique_lock<std::mutex> lk__(m->mu, std::defer_lock);
    if (!hvp_mpc_locked_by_this_thread) lk__.lock();
    HvpReentry re__;
    if (!m) return fail(-1, "mpc_solve: handle is NULL");
    if (batch < 0) return fail(-4, "mpc_solve: negative batch");
    if (batch == 0) return 0;
    if (!x0 || !mass || !params || !u || !x || !modes || !obj || !status || !nodes)
        return fail(-1, "mpc_solve: NULL array argument");
    hvp_ctx* c = m->ctx;
    const PmDev& S = m->S;
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t B = (size_t)batch, nl = S.nl, N = S.N;
    const size_t b_x0 = al256(B * nl * 16), b_m = al256(B * nl * 8), b_p = al256(B * S.npar * 8);
    const size_t b_fm = fixed_modes ? al256(B * nl * N * 4) : 0;
    const size_t b_u = al256(B * nl * N * 8), b_x = al256(B * nl * 2 * (N + 1) * 8);
    const size_t b_e = extra ? al256(B * (size_t)S.ne * 8 + 8) : 0, b_mo = al256(B * nl * N * 4);
    const size_t b_o = al256(B * 8), b_s = al256(B * 4);
    int rc = hvp_ensure_dbuf(c, b_x0 + b_m + b_p + b_fm + b_u + b_x + b_e + b_mo + b_o + 3 * b_s);
    if (rc) return rc;
    char* q = c->dbuf;
    double* dx0 = (double*)q; q += b_x0;
    double* dm = (double*)q; q += b_m;
    double* dp = (double*)q; q += b_p;
    int32_t* dfm = fixed_modes ? (int32_t*)q : nullptr; q += b_fm;
    double* du = (double*)q; q += b_u;
    double* dx = (double*)q; q += b_x;
    double* de = extra ? (double*)q : nullptr; q += b_e;
    int32_t* dmo = (int32_t*)q; q += b_mo;
    double* dob = (double*)q; q += b_o;
    int32_t* dst = (int32_t*)q; q += b_s;
    int32_t* dno = (int32_t*)q; q += b_s;
    int32_t* dit = qp_iters ? (int32_t*)q : nullptr;
    cudaStream_t st = c->stream;
    CUDA_TRY(cudaMemcpyAsync(dx0, x0, B * nl * 16, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dm, mass, B * nl * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dp, params, B * S.npar * 8, cudaMemcpyHostToDevice, st));
    if (fixed_modes) CUDA_TRY(cudaMemcpyAsync(dfm, fixed_modes, B * nl * N * 4, cudaMemcpyHostToDevice, st));
    rc = hvp_mpc_solve_dev(m, batch, dx0, dm, dp, dfm, du, dx, de, dmo, dob, dst, dno, dit, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(u, du, B * nl * N * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(x, dx, B * nl * 2 * (N + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (extra && S.ne) CUDA_TRY(cudaMemcpyAsync(extra, de, B * (size_t)S.ne * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(modes, dmo, B * nl * N * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(obj, dob, B * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(status, dst, B * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(nodes, dno, B * 4, cudaMemcpyDeviceToHost, st));
    if (qp_iters) CUDA_TRY(cudaMemcpyAsync(qp_iters, dit, B * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

// gathers the initial condition x0[b][i] = (xg[b][i][0][0], xg[b][i][1][0]) of an eval_cost guess
__global__ void pm_gather_x0(int64_t total, int np1, const double* __restrict__ xg, double* __restrict__ x0) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // t = (b * nl + i) * 2 + row
    if (t < total) x0[t] = xg[(size_t)t * np1];
}

extern "C" int hvp_mpc_eval_dev(hvp_mpc* m, int64_t batch, const double* mass, const double* params,
                                const double* xg, const double* ug, double* cost, void* stream) {
    if (!m) return fail(-1, "hvp_mpc_eval_dev: NULL handle");
    std::unique_lock<std::mutex> lk__(m->mu, std::defer_lock);
    if (!hvp_mpc_locked_by_this_thread) lk__.lock();
    HvpReentry re__;
    if (!m) return fail(-1, "mpc_eval: handle is NULL");
    if (batch < 0) return fail(-4, "mpc_eval: negative batch");
    if (batch == 0) return 0;
    if (!mass || !params || !xg || !ug || !cost) return fail(-1, "mpc_eval: NULL array argument");
    if (m->S.nel > 0) return fail(-4, "mpc_eval: not available for formulations with eliminated free copies");
    hvp_ctx* c = m->ctx;
    const PmDev& S = m->S;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    // scratch for the gathered initial conditions lives behind the handle (grow-only)
    const size_t need = (size_t)batch * S.nl * 2 * sizeof(double);
    if (need > m->x0cap) {
        if (m->x0buf) { CUDA_TRY(cudaStreamSynchronize(st)); CUDA_TRY(cudaFree(m->x0buf)); m->x0buf = nullptr; m->x0cap = 0; }
        CUDA_TRY(cudaMalloc(&m->x0buf, need + need / 2));
        m->x0cap = need + need / 2;
    }
    const int64_t total = batch * S.nl * 2;
    pm_gather_x0<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, S.N + 1, xg, m->x0buf);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(launch_pm_eval(S, batch, m->x0buf, mass, params, xg, ug, cost, st));
    c->launches += 2;
    return 0;
}

extern "C" int hvp_mpc_eval_host(hvp_mpc* m, int64_t batch, const double* mass, const double* params,
                                 const double* xg, const double* ug, double* cost) {
    if (!m) return fail(-1, "hvp_mpc_eval_host: NULL handle");
    std::unique_lock<std::mutex> lk__(m->mu, std::defer_lock);
    if (!hvp_mpc_locked_by_this_thread) lk__.lock();
    HvpReentry re__;
    if (!m) return fail(-1, "mpc_eval: handle is NULL");
    if (batch < 0) return fail(-4, "mpc_eval: negative batch");
    if (batch == 0) return 0;
    if (!mass || !params || !xg || !ug || !cost) return fail(-1, "mpc_eval: NULL array argument");
    hvp_ctx* c = m->ctx;
    const PmDev& S = m->S;
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t B = (size_t)batch, nl = S.nl, N = S.N;
    const size_t b_m = al256(B * nl * 8), b_p = al256(B * S.npar * 8), b_x = al256(B * nl * 2 * (N + 1) * 8);
    const size_t b_u = al256(B * nl * N * 8), b_c = al256(B * 8);
    int rc = hvp_ensure_dbuf(c, b_m + b_p + b_x + b_u + b_c);
    if (rc) return rc;
    char* q = c->dbuf;
    double* dm = (double*)q; q += b_m;
    double* dp = (double*)q; q += b_p;
    double* dx = (double*)q; q += b_x;
    double* du = (double*)q; q += b_u;
    double* dc = (double*)q;
    cudaStream_t st = c->stream;
    CUDA_TRY(cudaMemcpyAsync(dm, mass, B * nl * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dp, params, B * S.npar * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dx, xg, B * nl * 2 * (N + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(du, ug, B * nl * N * 8, cudaMemcpyHostToDevice, st));
    rc = hvp_mpc_eval_dev(m, batch, dm, dp, dx, du, dc, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(cost, dc, B * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}
