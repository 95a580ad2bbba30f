// coord.cu -- fused coordinator glue of the switching ("g") ADMM controller.
//
// One consensus round of fleet_g_admm.py:255-301 (+ the restated round logic of dmpcpwa's GAdmmCoordinator /
// dmpcrl's MpcAdmm, hybrid_vehicle_platoon_b200/fleet_g_admm.py) is, for every agent: solve the fixed-sequence QP
// (compiled-MPC kernel, one launch per role), then
//     z_j   <- mean of the copies of vehicle j's trajectory (own, the copy held by j+1, the copy held by j-1)
//     y     <- y + rho (copy - z)                         for the agent's own block and its neighbour copies
//     u_j   <- the QP's inputs where the QP was solved
//     seq_j <- PWA region sequence of the roll-out of u_j from the current state (the next round's fixed modes)
// and the next round's parameter vector [leader window | y blocks | z blocks].  As torch ops this was ~280 tiny
// launches per round (r02i launch list: 1.1 ms of a 2.1 ms round at 1024 scenarios x 15 vehicles); here it is ONE
// kernel, one thread per (scenario, vehicle): a vehicle's update needs only its neighbours' QP outputs, which it reads
// straight from the role buffers, recomputing the two neighbouring z's instead of synchronising.
//
// Arithmetic: every expression is evaluated in the order of the torch formulation with explicit round-to-nearest
// multiplies and adds (no FMA contraction), so the consensus variables are bit-identical to the unfused path
// (tests/test_gpu_sweep.py compares the two); only the per-scenario cost is summed in vehicle order.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/hvp.h"
#include "hvp_internal.h"

namespace hvp {

struct GRole {                    // one role = one compiled formulation = one launch (leader / last / interior)
    double* params;               // [count*S][npar]   next round's parameters (written here)
    double* x0;                   // [count*S][2]
    double* mass;                 // [count*S]
    int32_t* fm;                  // [count*S][N]      next round's fixed mode sequences
    const double* uo;             // [count*S][N]      QP outputs of the round just solved
    const double* xo;             // [count*S][2][N+1]
    const double* eo;             // [count*S][ne]     copies: front (if any) then back (if any), (2, N+1) each
    const double* ob;             // [count*S]
    const int32_t* st;            // [count*S]
    int first, count, npar, ne;
};

struct GArgs {
    int n, N, S, init;
    double rho, mug;
    double edge[6], cf[7], bg[7], dd[7];
    GRole role[3];                // 0: vehicle 0, 1: vehicle n-1, 2: vehicles 1..n-2 (count 0 when n == 2)
    const double* x;              // [S][n][2]    current state
    const double* mass;           // [S][n]
    const double* lwin;           // [S][2][N+1]  leader window
    double* y;                    // [S][n][3][2][N+1]  multipliers of the blocks (front copy, own, back copy)
    double* u;                    // [S][n][N]    current inputs (in: start / previous round; out: updated)
    double* tr;                   // [S][n][2][N+1]  PWA roll-out of u
    double* cost;                 // [S]
    uint8_t* ok;                  // [S]  and-accumulated over the rounds of a warm start (set to 1 by init)
};

__device__ __forceinline__ int role_of(const GArgs& A, int i) { return i == 0 ? 0 : (i == A.n - 1 ? 1 : 2); }

// QP outputs of vehicle j in scenario s
struct VOut {
    const double *uo, *xo, *cf, *cb;
    bool good;
    double ob;
};
__device__ __forceinline__ VOut vout(const GArgs& A, int s, int j) {
    const GRole& R = A.role[role_of(A, j)];
    const int64_t b = (int64_t)s * R.count + (j - R.first);
    const int np1 = A.N + 1;
    VOut o;
    o.uo = R.uo + b * A.N;
    o.xo = R.xo + b * 2 * np1;
    const double* e = R.eo + b * R.ne;
    o.cf = j > 0 ? e : nullptr;
    o.cb = j < A.n - 1 ? e + (j > 0 ? 2 * np1 : 0) : nullptr;
    o.good = R.st[b] == 2;
    o.ob = R.ob[b];
    return o;
}

// z of vehicle j, entry e of its (2, N+1) block: (own + copy held by j+1) + copy held by j-1, over the count
__device__ __forceinline__ double zval(const GArgs& A, int s, int j, int e) {
    const VOut me = vout(A, s, j);
    double acc = me.xo[e];
    double cnt = 1.0;
    if (j < A.n - 1) { acc = __dadd_rn(acc, vout(A, s, j + 1).cf[e]); cnt += 1.0; }
    if (j > 0) { acc = __dadd_rn(acc, vout(A, s, j - 1).cb[e]); cnt += 1.0; }
    return acc / cnt;
}

__global__ void __launch_bounds__(128)
gadmm_round_kernel(const __grid_constant__ GArgs A) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = A.n, N = A.N, np1 = N + 1, blk = 2 * np1;
    if (t >= (int64_t)A.S * n) return;
    const int s = (int)(t / n), j = (int)(t - (int64_t)s * n);
    const GRole& R = A.role[role_of(A, j)];
    const int64_t b = (int64_t)s * R.count + (j - R.first);
    double* yb = A.y + (size_t)t * 3 * blk;          // [front | own | back]
    double* uj = A.u + (size_t)t * N;
    double* trj = A.tr + (size_t)t * blk;
    const double m = A.mass[t];
    const double p0 = A.x[2 * t], v0 = A.x[2 * t + 1];

    if (!A.init) {
        // ---- inputs of the solved QPs, multiplier update with the new z ----
        const VOut me = vout(A, s, j);
        if (me.good)
            for (int k = 0; k < N; ++k) uj[k] = me.uo[k];
        for (int e = 0; e < blk; ++e) {
            const double zj = zval(A, s, j, e);
            yb[blk + e] = __dadd_rn(yb[blk + e], __dmul_rn(A.rho, __dadd_rn(me.xo[e], -zj)));
            if (j > 0) yb[e] = __dadd_rn(yb[e], __dmul_rn(A.rho, __dadd_rn(me.cf[e], -zval(A, s, j - 1, e))));
            if (j < n - 1)
                yb[2 * blk + e] = __dadd_rn(yb[2 * blk + e], __dmul_rn(A.rho, __dadd_rn(me.cb[e], -zval(A, s, j + 1, e))));
        }
    } else {
        for (int e = 0; e < 3 * blk; ++e) yb[e] = 0.0;
    }
    // ---- PWA roll-out of u_j from the current state: trajectory and region sequence ----
    {
        double p = p0, v = v0;
        trj[0] = p; trj[np1] = v;
        for (int k = 0; k < N; ++k) {
            int r = 0;
            for (int q = 0; q < 6; ++q) r += (A.edge[q] + 1e-9 < v) ? 1 : 0;      // bucketize(v, edges + 1e-9, right=False)
            R.fm[b * N + k] = r;
            const double a = __dadd_rn(1.0, (-A.cf[r]) / m);
            const double bb = A.bg[r] / m;
            const double c = __dadd_rn(-A.mug, -(A.dd[r] / m));
            const double pn = __dadd_rn(p, v);
            const double vn = __dadd_rn(__dadd_rn(__dmul_rn(a, v), __dmul_rn(bb, uj[k])), c);
            p = pn; v = vn;
            trj[k + 1] = p; trj[np1 + k + 1] = v;
        }
    }
    // ---- next round's parameters: [leader window | y blocks | z blocks], blocks = [front (j > 0), own, back (j < n-1)] ----
    double* P = R.params + (size_t)b * R.npar;
    R.x0[2 * b] = p0; R.x0[2 * b + 1] = v0;
    R.mass[b] = m;
    const double* lw = A.lwin + (size_t)s * blk;
    for (int e = 0; e < blk; ++e) P[e] = lw[e];
    int o = blk;
    if (j > 0) { for (int e = 0; e < blk; ++e) P[o + e] = yb[e]; o += blk; }
    for (int e = 0; e < blk; ++e) P[o + e] = yb[blk + e];
    o += blk;
    if (j < n - 1) { for (int e = 0; e < blk; ++e) P[o + e] = yb[2 * blk + e]; o += blk; }
    // z blocks.  init: z = the roll-out of the start inputs -- the neighbours' roll-outs are recomputed here (cheap)
    for (int d = -1; d <= 1; ++d) {
        const int jj = j + d;
        if (jj < 0 || jj >= n) continue;
        if (!A.init) {
            for (int e = 0; e < blk; ++e) P[o + e] = zval(A, s, jj, e);
        } else if (d == 0) {
            for (int e = 0; e < blk; ++e) P[o + e] = trj[e];
        } else {
            const int64_t tt = t + d;
            const double mm = A.mass[tt];
            const double* uu = A.u + (size_t)tt * N;
            double p = A.x[2 * tt], v = A.x[2 * tt + 1];
            P[o] = p; P[o + np1] = v;
            for (int k = 0; k < N; ++k) {
                int r = 0;
                for (int q = 0; q < 6; ++q) r += (A.edge[q] + 1e-9 < v) ? 1 : 0;
                const double a = __dadd_rn(1.0, (-A.cf[r]) / mm);
                const double bb = A.bg[r] / mm;
                const double c = __dadd_rn(-A.mug, -(A.dd[r] / mm));
                const double pn = __dadd_rn(p, v);
                const double vn = __dadd_rn(__dadd_rn(__dmul_rn(a, v), __dmul_rn(bb, uu[k])), c);
                p = pn; v = vn;
                P[o + k + 1] = p; P[o + np1 + k + 1] = v;
            }
        }
        o += blk;
    }
    // ---- per scenario: cost of the round (solved QPs only, vehicle order) and the and-accumulated status ----
    if (j == 0) {
        if (A.init) {
            A.ok[s] = 1; A.cost[s] = 0.0;
        } else {
            double c = 0.0;
            bool all = true;
            for (int i = 0; i < n; ++i) {
                const VOut vo = vout(A, s, i);
                if (vo.good) c = __dadd_rn(c, vo.ob); else all = false;
            }
            A.cost[s] = c;
            if (!all) A.ok[s] = 0;
        }
    }
}


// ---- decentralized / sequential controllers: constant-velocity extrapolation of every vehicle's neighbours and the
// leader-trajectory window of the timestep (fleet_decent_mld.py:329-331, :421-428) in one launch -------------------------
struct ObsArgs {
    int n, N, S, leader_index;
    long long lx_len, lx_sstride;     // leader trajectory: [S or 1][2][lx_len], scenario stride 0 when shared
    double ts;
    const double* x;                  // [S][n][2]
    const double* lx;
    const long long* t;               // device-side timestep index
    double *xf, *xb, *xl;             // [S][n][2][N+1] each
};
__global__ void __launch_bounds__(128)
decent_observe_kernel(const __grid_constant__ ObsArgs A) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = A.n, N = A.N, np1 = N + 1, blk = 2 * np1;
    if (tid >= (int64_t)A.S * n) return;
    const int s = (int)(tid / n), i = (int)(tid - (int64_t)s * n);
    // my own prediction is my back neighbour's x_front and my front neighbour's x_back
    double p = A.x[2 * tid];
    const double v = A.x[2 * tid + 1];
    double* f = i + 1 < n ? A.xf + (size_t)(tid + 1) * blk : nullptr;
    double* b = i > 0 ? A.xb + (size_t)(tid - 1) * blk : nullptr;
    for (int k = 0; k <= N; ++k) {
        if (f) { f[k] = p; f[np1 + k] = v; }
        if (b) { b[k] = p; b[np1 + k] = v; }
        p = __dadd_rn(p, __dmul_rn(A.ts, v));            // sequential sums, as the reference (p_{k+1} = p_k + ts v_k)
    }
    if (i == A.leader_index) {
        const long long t = *A.t;
        const double* l = A.lx + (size_t)s * A.lx_sstride;
        double* o = A.xl + (size_t)tid * blk;
        for (int k = 0; k <= N; ++k) { o[k] = l[t + k]; o[np1 + k] = l[A.lx_len + t + k]; }
    }
}

// ---- naive ADMM: z- and y-update of a consensus round and the next round's parameter vectors
// (fleet_naive_admm.py:421-468) in one launch, read straight from the role solves' outputs -------------------------------
struct ARole {
    double* params;                   // [count*S][5 * 2(N+1)]: leader window | y_front | z_front | y_back | z_back
    const double* xo;                 // [count*S][2][N+1]   own prediction of the round just solved
    const double* eo;                 // [count*S][ne]       copies: x_front (if not front) then x_back (if not trailer)
    int count, ne, has_front, has_back;
};
struct AdmmArgs {
    int n, N, S, nroles, pack_only;
    double rho;
    ARole role[4];
    int role_of[64], local_of[64];    // vehicle -> role, index among the role's vehicles
    const double* lwin;               // [S][2][N+1]
    double *y_front, *y_back, *zf, *zb, *xs;   // [S][n][2][N+1] each
};
struct AOut { const double *own, *cf, *cb; };
__device__ __forceinline__ AOut aout(const AdmmArgs& A, int s, int j) {
    const ARole& R = A.role[A.role_of[j]];
    const int64_t b = (int64_t)s * R.count + A.local_of[j];
    const int blk = 2 * (A.N + 1);
    AOut o;
    o.own = R.xo + b * blk;
    const double* e = R.eo + b * R.ne;
    o.cf = R.has_front ? e : nullptr;
    o.cb = R.has_back ? e + (R.has_front ? blk : 0) : nullptr;
    return o;
}
// z of vehicle j: (own + x_front copy held by j+1) / 2 at the front, (own + x_back copy held by j-1) / 2 at the rear,
// (own + both) / 3 inside -- in exactly this order of additions (fleet_naive_admm.py:421-447)
__device__ __forceinline__ double azval(const AdmmArgs& A, int s, int j, int e) {
    const double own = aout(A, s, j).own[e];
    if (j == 0) return __dadd_rn(own, aout(A, s, 1).cf[e]) / 2.0;
    if (j == A.n - 1) return __dadd_rn(own, aout(A, s, j - 1).cb[e]) / 2.0;
    return __dadd_rn(__dadd_rn(own, aout(A, s, j + 1).cf[e]), aout(A, s, j - 1).cb[e]) / 3.0;
}
__global__ void __launch_bounds__(128)
admm_round_kernel(const __grid_constant__ AdmmArgs A) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = A.n, blk = 2 * (A.N + 1);
    if (tid >= (int64_t)A.S * n) return;
    const int s = (int)(tid / n), j = (int)(tid - (int64_t)s * n);
    const AOut me = aout(A, s, j);
    const ARole& R = A.role[A.role_of[j]];
    double* P = R.params + ((size_t)s * R.count + A.local_of[j]) * (5 * blk);
    const size_t o = (size_t)tid * blk;
    const double* lw = A.lwin + (size_t)s * blk;
    for (int e = 0; e < blk; ++e) {
        double yf = A.y_front[o + e], yb = A.y_back[o + e], zfv = A.zf[o + e], zbv = A.zb[o + e];
        if (A.pack_only) {
            P[e] = lw[e]; P[blk + e] = yf; P[2 * blk + e] = zfv; P[3 * blk + e] = yb; P[4 * blk + e] = zbv;
            continue;
        }
        A.xs[o + e] = me.own[e];
        if (j > 0) {                  // y_front += rho (x_front copy - z of the vehicle in front); next target z_{j-1}
            zfv = azval(A, s, j - 1, e);
            yf = __dadd_rn(yf, __dmul_rn(A.rho, __dadd_rn(me.cf[e], -zfv)));
        }
        if (j < n - 1) {
            zbv = azval(A, s, j + 1, e);
            yb = __dadd_rn(yb, __dmul_rn(A.rho, __dadd_rn(me.cb[e], -zbv)));
        }
        A.y_front[o + e] = yf; A.y_back[o + e] = yb; A.zf[o + e] = zfv; A.zb[o + e] = zbv;
        P[e] = lw[e]; P[blk + e] = yf; P[2 * blk + e] = zfv; P[3 * blk + e] = yb; P[4 * blk + e] = zbv;
    }
}

}  // namespace hvp

using namespace hvp;

// hvp.h: hvp_gadmm_round_dev
extern "C" int hvp_gadmm_round_dev(hvp_ctx* c, const hvp_gadmm_round* g, void* stream) {
    if (!c || !g) return hvp_fail(-1, "gadmm_round: NULL argument");
    if (g->n < 2 || g->N < 1 || g->N > 16 || g->S < 0) return hvp_fail(-4, "gadmm_round: bad sizes (n=%d N=%d S=%d)", g->n, g->N, g->S);
    if (g->S == 0) return 0;
    if (!g->x || !g->mass || !g->lwin || !g->y || !g->u || !g->tr || !g->cost || !g->ok)
        return hvp_fail(-1, "gadmm_round: NULL array argument");
    GArgs A;
    A.n = g->n; A.N = g->N; A.S = g->S; A.init = g->init; A.rho = g->rho;
    A.mug = g->mug;
    for (int r = 0; r < 7; ++r) { A.cf[r] = g->cf[r]; A.bg[r] = g->bg[r]; A.dd[r] = g->dd[r]; }
    for (int q = 0; q < 6; ++q) A.edge[q] = g->edge[q];
    const int np1 = g->N + 1;
    for (int r = 0; r < 3; ++r) {
        const hvp_gadmm_role& s = g->role[r];
        GRole& R = A.role[r];
        R.first = r == 0 ? 0 : (r == 1 ? g->n - 1 : 1);
        R.count = r == 2 ? g->n - 2 : 1;
        const int na = (r == 2) ? 3 : 2;
        R.npar = (1 + 2 * na) * 2 * np1;
        R.ne = (na - 1) * 2 * np1;
        R.params = s.params; R.x0 = s.x0; R.mass = s.mass; R.fm = s.fixed_modes;
        R.uo = s.u; R.xo = s.x; R.eo = s.extra; R.ob = s.obj; R.st = s.status;
        if (R.count > 0 && (!R.params || !R.x0 || !R.mass || !R.fm || !R.uo || !R.xo || !R.eo || !R.ob || !R.st))
            return hvp_fail(-1, "gadmm_round: NULL role buffer (role %d)", r);
    }
    A.x = g->x; A.mass = g->mass; A.lwin = g->lwin; A.y = g->y; A.u = g->u; A.tr = g->tr; A.cost = g->cost; A.ok = g->ok;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return hvp_fail(-100 - (int)e, "gadmm_round: cudaSetDevice: %s", cudaGetErrorString(e));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const int64_t threads = (int64_t)g->S * g->n;
    gadmm_round_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(A);
    e = cudaGetLastError();
    if (e != cudaSuccess) return hvp_fail(-100 - (int)e, "gadmm_round: launch failed: %s", cudaGetErrorString(e));
    c->launches += 1;
    return 0;
}

// hvp.h: hvp_decent_observe_dev
extern "C" int hvp_decent_observe_dev(hvp_ctx* c, const hvp_decent_observe* g, void* stream) {
    if (!c || !g) return hvp_fail(-1, "decent_observe: NULL argument");
    if (g->n < 1 || g->N < 1 || g->S < 0 || g->leader_index < 0 || g->leader_index >= g->n)
        return hvp_fail(-4, "decent_observe: bad sizes (n=%d N=%d S=%d leader=%d)", g->n, g->N, g->S, g->leader_index);
    if (g->S == 0) return 0;
    if (!g->x || !g->leader_x || !g->t || !g->xf || !g->xb || !g->xl) return hvp_fail(-1, "decent_observe: NULL array argument");
    ObsArgs A;
    A.n = g->n; A.N = g->N; A.S = g->S; A.leader_index = g->leader_index; A.ts = g->ts;
    A.lx_len = g->leader_len; A.lx_sstride = g->leader_per_scenario ? 2 * g->leader_len : 0;
    A.x = g->x; A.lx = g->leader_x; A.t = (const long long*)g->t; A.xf = g->xf; A.xb = g->xb; A.xl = g->xl;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return hvp_fail(-100 - (int)e, "decent_observe: cudaSetDevice: %s", cudaGetErrorString(e));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const int64_t threads = (int64_t)g->S * g->n;
    decent_observe_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(A);
    e = cudaGetLastError();
    if (e != cudaSuccess) return hvp_fail(-100 - (int)e, "decent_observe: launch failed: %s", cudaGetErrorString(e));
    c->launches += 1;
    return 0;
}

// hvp.h: hvp_admm_round_dev
extern "C" int hvp_admm_round_dev(hvp_ctx* c, const hvp_admm_round* g, void* stream) {
    if (!c || !g) return hvp_fail(-1, "admm_round: NULL argument");
    if (g->n < 2 || g->n > 64 || g->N < 1 || g->S < 0 || g->nroles < 1 || g->nroles > 4)
        return hvp_fail(-4, "admm_round: bad sizes (n=%d N=%d S=%d roles=%d)", g->n, g->N, g->S, g->nroles);
    if (g->S == 0) return 0;
    if (!g->lwin || !g->y_front || !g->y_back || !g->zf || !g->zb || !g->xs) return hvp_fail(-1, "admm_round: NULL array argument");
    AdmmArgs A;
    A.n = g->n; A.N = g->N; A.S = g->S; A.nroles = g->nroles; A.rho = g->rho; A.pack_only = g->pack_only;
    const int blk = 2 * (g->N + 1);
    int counts[4] = {0, 0, 0, 0};
    for (int i = 0; i < g->n; ++i) {
        const int r = g->role_of[i];
        if (r < 0 || r >= g->nroles) return hvp_fail(-4, "admm_round: role_of[%d] = %d out of range", i, r);
        A.role_of[i] = r; A.local_of[i] = counts[r]++;
    }
    for (int r = 0; r < g->nroles; ++r) {
        const hvp_admm_role& s = g->role[r];
        ARole& R = A.role[r];
        R.params = s.params; R.xo = s.x; R.eo = s.extra; R.count = counts[r];
        R.has_front = s.has_front; R.has_back = s.has_back; R.ne = (s.has_front + s.has_back) * blk;
        if (counts[r] > 0 && (!R.params || !R.xo || (R.ne > 0 && !R.eo))) return hvp_fail(-1, "admm_round: NULL role buffer (role %d)", r);
    }
    // a vehicle reads the copies its neighbours hold of it: vehicle i > 0 must hold x_front, i < n-1 must hold x_back
    for (int i = 0; i < g->n; ++i) {
        const ARole& R = A.role[A.role_of[i]];
        if ((i > 0) != (R.has_front != 0) || (i < g->n - 1) != (R.has_back != 0))
            return hvp_fail(-4, "admm_round: role of vehicle %d does not match its position (front / trailer copies)", i);
    }
    A.lwin = g->lwin; A.y_front = g->y_front; A.y_back = g->y_back; A.zf = g->zf; A.zb = g->zb; A.xs = g->xs;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return hvp_fail(-100 - (int)e, "admm_round: cudaSetDevice: %s", cudaGetErrorString(e));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const int64_t threads = (int64_t)g->S * g->n;
    admm_round_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(A);
    e = cudaGetLastError();
    if (e != cudaSuccess) return hvp_fail(-100 - (int)e, "admm_round: launch failed: %s", cudaGetErrorString(e));
    c->launches += 1;
    return 0;
}
