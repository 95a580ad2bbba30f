/*
 * hvp_oracle_mpc.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE), second part:
 * the multi-vehicle / ADMM / discrete-gear MPC formulations of the reference.
 *
 * PARITY STATUS: "parity unpinned" (see hvp_oracle.c): the reference solves these models with
 * Gurobi behind the un-vendored dmpcpwa==0.0.2; this file restates the problem DEFINITIONS
 *   CENT   mpcs/cent_mld.py:48-177            (MpcMldCent on MpcMldCentDecup)
 *   LOCAL  fleet_decent_mld.py:61-208, fleet_seq_mld.py:63-219
 *   EVENT  fleet_event_based.py:72-291
 *   ADMM   fleet_naive_admm.py:63-237
 *   GADMM  fleet_g_admm.py:55-158 (+ dmpcrl MpcAdmm consensus terms, UNVERIFIED-3P)
 * on the PWA-gear model (models.py:397-492) or the PWA-friction model with the discrete gears of
 * MpcGear.setup_gears (models.py:288-332, mpcs/mpc_gear.py:30-114), and computes the global
 * optimum as the minimum over mode sequences of the fixed-mode convex QP:
 * exhaustive reachability-pruned enumeration (ground truth, small sizes) or depth-first branch and
 * bound whose node bound relaxes the not-yet-fixed stages to free velocity increments.
 *
 * Deliberately different from the CUDA product path: INPUT space (variables u_{i,k}), dense
 * condensed rows, the generic dense QP solver of hvp_oracle.c; the formulation is typed from the
 * reference files independently of csrc/pm_build.cu.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int hvo_qp_solve(int n, int m, const double* H, const double* g, double c0, const double* A,
                 const double* b, const double* wmax, double* x, double* lam_out, double* obj,
                 int* iters_out);
void hvo_pwa_gear_system(double m, double ts, double a[7], double b[7], double c[7], double lo[7],
                         double hi[7]);

#define MAXMODES 12
#define MAXN 16
#define MAXVEH 16
#define MAXVAR 160
#define MAXROWS 2400

enum { K_CENT = 1, K_LOCAL = 2, K_EVENT = 3, K_ADMM = 4, K_GADMM = 5 };
enum { F_FRONT = 1, F_LEADER = 2, F_TRAILER = 4, F_REAL_REF = 8 };
#define NO_LEADER (-100)

static const double QX_P = 1.0, QX_V = 0.1, Q_U = 1.0, W_SLACK = 1e4;
static const double A_ACC = 2.5, A_DEC = -2.0, D_SAFE = 25.0;
static const double V_MIN = 3.94, V_MAX = 45.84, P_MIN = 0.0, P_MAX = 10000.0, U_MIN = -1.0, U_MAX = 1.0;
static const double C_FRIC = 0.5, MU = 0.01, GRAV = 9.8;
static const double BGEAR[6] = {4057, 2945, 2116, 1607, 1166, 838};
static const double VL[6] = {3.94, 5.43, 7.56, 9.96, 13.70, 19.10};
static const double VH[6] = {9.46, 13.04, 18.15, 23.90, 32.93, 45.84};

/* mode table of one vehicle: v+ = a v + b u + c on lo <= v <= hi, ts = 1.
 * model 0: PWA-gear (7 regions).  model 1: PWA friction (2 regions, models.py:288-332, forward
 * Euler models.py:370-387) combined with the six gears of MpcGear (mpc_gear.py:80-110): mode
 * (f, j) has B = b_j / m acting on the throttle u_g and is valid where friction region f and the
 * gear window [vl_j, vh_j] overlap. */
int hvo_mode_table(int model, double m, double* a, double* b, double* c, double* lo, double* hi, int* gear) {
    if (model == 0) {
        hvo_pwa_gear_system(m, 1.0, a, b, c, lo, hi);
        static const int g[7] = {1, 2, 3, 4, 4, 5, 6};
        for (int r = 0; r < 7; ++r) gear[r] = g[r];
        return 7;
    }
    double beta = (3 * C_FRIC * V_MAX * V_MAX) / 16, alpha = V_MAX / 2;
    double c1 = beta / alpha, c2 = (C_FRIC * V_MAX * V_MAX - beta) / (V_MAX - alpha);
    double d = beta - alpha * ((C_FRIC * V_MAX * V_MAX - beta) / (V_MAX - alpha));
    for (int f = 0; f < 2; ++f)
        for (int j = 0; j < 6; ++j) {
            int r = f * 6 + j;
            a[r] = 1.0 + (-(f == 0 ? c1 : c2) / m);
            b[r] = BGEAR[j] * (1.0 / m);
            c[r] = f == 0 ? (-MU * GRAV) : (-MU * GRAV - d / m);
            lo[r] = fmax(f == 0 ? -INFINITY : alpha, VL[j]);
            hi[r] = fmin(f == 0 ? alpha : INFINITY, VH[j]);
            gear[r] = j + 1;
        }
    return 12;
}

typedef struct {
    int kind, model, nl, N, flags, leader_index, n_front, n_behind;
    double d0, t0, tight, rho;
    const double* x0;      /* [nl][2] */
    const double* mass;    /* [nl] */
    const double* params;  /* blocks of (2, N+1), see include/hvp.h */
    int ne, nvar;          /* extras, total variables */
    int R;
    double a[MAXVEH][MAXMODES], b[MAXVEH][MAXMODES], c[MAXVEH][MAXMODES], lo[MAXMODES], hi[MAXMODES];
    int gear[MAXMODES];
} mpc_prob;

typedef struct { double c; double g[MAXVAR]; } aff;

typedef struct {
    const mpc_prob* P;
    int n;                              /* variables */
    aff p[MAXVEH][MAXN + 1], v[MAXVEH][MAXN + 1];
    double *H, *g, c0, *A, *b, *w;
    int m;
} qp_build;

static aff aff_const(double c) { aff e; memset(&e, 0, sizeof e); e.c = c; return e; }
static aff aff_var(int j) { aff e; memset(&e, 0, sizeof e); e.g[j] = 1.0; return e; }
static aff aff_axpy(aff x, double s, const aff* y, int n) {
    x.c += s * y->c;
    for (int j = 0; j < n; ++j) x.g[j] += s * y->g[j];
    return x;
}
static aff par(const mpc_prob* P, int block, int row, int k) {
    return aff_const(P->params[block * 2 * (P->N + 1) + row * (P->N + 1) + k]);
}

static void add_res(qp_build* Q, const aff* e, double wgt) {
    int n = Q->n;
    for (int i = 0; i < n; ++i) {
        if (e->g[i] == 0.0) continue;
        for (int j = 0; j < n; ++j) Q->H[i * n + j] += 2 * wgt * e->g[i] * e->g[j];
        Q->g[i] += 2 * wgt * e->c * e->g[i];
    }
    Q->c0 += wgt * e->c * e->c;
}
static void add_lin(qp_build* Q, double coef, const aff* e) {
    for (int i = 0; i < Q->n; ++i) Q->g[i] += coef * e->g[i];
    Q->c0 += coef * e->c;
}
/* e <= 0, hard (w = inf) or L1-penalised with weight w.  Rows without variables are resolved here:
 * returns 1 if a hard constant row is violated. */
static int add_row(qp_build* Q, const aff* e, double w) {
    int n = Q->n, any = 0;
    for (int j = 0; j < n; ++j) if (e->g[j] != 0.0) any = 1;
    if (!any) {
        if (e->c > 0) {
            if (isfinite(w)) Q->c0 += w * e->c;
            else if (e->c > 1e-9) return 1;
        }
        return 0;
    }
    for (int j = 0; j < n; ++j) Q->A[(size_t)Q->m * n + j] = e->g[j];
    Q->b[Q->m] = -e->c; Q->w[Q->m] = w;
    Q->m++;
    return 0;
}

/* ||x - y - sigma(x)||^2_Qx, sigma(x) = [-t0 v - d0, 0]  (spacing_policy.py:13-37) */
static void add_track(qp_build* Q, const aff* xp, const aff* xv, const aff* yp, const aff* yv) {
    const mpc_prob* P = Q->P;
    aff e = *xp;
    e = aff_axpy(e, -1.0, yp, Q->n);
    e = aff_axpy(e, P->t0, xv, Q->n);
    e.c += P->d0;
    add_res(Q, &e, QX_P);
    aff f = *xv;
    f = aff_axpy(f, -1.0, yv, Q->n);
    add_res(Q, &f, QX_V);
}
static void add_pair(qp_build* Q, const aff* xp, const aff* xv, const aff* yp, const aff* yv) {  /* ||x - y||^2_Qx */
    aff e = aff_axpy(*xp, -1.0, yp, Q->n);
    add_res(Q, &e, QX_P);
    aff f = aff_axpy(*xv, -1.0, yv, Q->n);
    add_res(Q, &f, QX_V);
}
static int add_safe(qp_build* Q, const aff* behind_p, const aff* ahead_p) {  /* p_behind <= p_ahead - d_safe + s */
    aff e = aff_axpy(*behind_p, -1.0, ahead_p, Q->n);
    e.c += D_SAFE;
    return add_row(Q, &e, W_SLACK);
}

/* Dense QP of the problem for the mode assignment modes[i*N+k] (-1 = stage relaxed to a free
 * velocity increment: used by the branch-and-bound bound only).  Returns 1 if trivially infeasible. */
static int build_qp(const mpc_prob* P, const int* modes, qp_build* Q) {
    const int nl = P->nl, N = P->N, n = P->nvar;
    Q->P = P; Q->n = n; Q->m = 0; Q->c0 = 0;
    memset(Q->H, 0, sizeof(double) * n * n);
    memset(Q->g, 0, sizeof(double) * n);
    /* MLD dynamics with the mode fixed: condensed affine state maps */
    for (int i = 0; i < nl; ++i) {
        Q->p[i][0] = aff_const(P->x0[2 * i]);
        Q->v[i][0] = aff_const(P->x0[2 * i + 1]);
        for (int k = 0; k < N; ++k) {
            int r = modes[i * N + k];
            aff u = aff_var(i * N + k);
            Q->p[i][k + 1] = aff_axpy(Q->p[i][k], 1.0, &Q->v[i][k], n);
            if (r >= 0) {
                aff nv = aff_const(P->c[i][r]);
                nv = aff_axpy(nv, P->a[i][r], &Q->v[i][k], n);
                nv = aff_axpy(nv, P->b[i][r], &u, n);
                Q->v[i][k + 1] = nv;
            } else {
                Q->v[i][k + 1] = aff_axpy(Q->v[i][k], 1.0, &u, n);
            }
        }
    }
    int bad = 0;
    for (int i = 0; i < nl; ++i) {
        for (int k = 0; k < N; ++k) {
            int r = modes[i * N + k];
            aff u = aff_var(i * N + k);
            if (r >= 0) {
                add_res(Q, &u, Q_U);                                   /* control effort */
                aff e;                                                 /* region rows S x <= T */
                if (isfinite(P->hi[r])) { e = Q->v[i][k]; e.c -= P->hi[r]; bad |= add_row(Q, &e, INFINITY); }
                if (isfinite(P->lo[r])) { e = aff_axpy(aff_const(P->lo[r]), -1.0, &Q->v[i][k], n); bad |= add_row(Q, &e, INFINITY); }
                e = u; e.c -= U_MAX; bad |= add_row(Q, &e, INFINITY);  /* F u <= G */
                e = aff_axpy(aff_const(U_MIN), -1.0, &u, n); bad |= add_row(Q, &e, INFINITY);
            }
            /* accel rows: a_dec <= v+ - v - k tight ; v+ - v <= a_acc - k tight */
            aff dv = aff_axpy(Q->v[i][k + 1], -1.0, &Q->v[i][k], n);
            aff e = dv; e.c -= A_ACC - k * P->tight; bad |= add_row(Q, &e, INFINITY);
            e = aff_axpy(aff_const(A_DEC + k * P->tight), -1.0, &dv, n); bad |= add_row(Q, &e, INFINITY);
        }
        for (int k = 1; k <= N; ++k) {                                  /* D x <= E, k = 1..N */
            aff e = Q->p[i][k]; e.c -= P_MAX; bad |= add_row(Q, &e, INFINITY);
            e = aff_axpy(aff_const(P_MIN), -1.0, &Q->p[i][k], n); bad |= add_row(Q, &e, INFINITY);
            e = Q->v[i][k]; e.c -= V_MAX; bad |= add_row(Q, &e, INFINITY);
            e = aff_axpy(aff_const(V_MIN), -1.0, &Q->v[i][k], n); bad |= add_row(Q, &e, INFINITY);
        }
    }
    const int front = P->flags & F_FRONT, leader = P->flags & F_LEADER, trailer = P->flags & F_TRAILER;
    const int real_ref = P->flags & F_REAL_REF;
    const int np1 = N + 1, X = nl * N;     /* first extra variable */
    switch (P->kind) {
        case K_CENT: {                     /* cent_mld.py:81-177 */
            int L = P->leader_index;
            for (int k = 0; k <= N; ++k) {
                aff rp = par(P, 0, 0, k), rv = par(P, 0, 1, k);
                if (!real_ref) add_pair(Q, &Q->p[L][k], &Q->v[L][k], &rp, &rv);
                else { add_track(Q, &Q->p[0][k], &Q->v[0][k], &rp, &rv); bad |= add_safe(Q, &Q->p[0][k], &rp); }
                for (int i = 1; i < nl; ++i) {
                    add_track(Q, &Q->p[i][k], &Q->v[i][k], &Q->p[i - 1][k], &Q->v[i - 1][k]);
                    bad |= add_safe(Q, &Q->p[i][k], &Q->p[i - 1][k]);
                }
            }
        } break;
        case K_LOCAL: {                    /* fleet_decent_mld.py:108-208, fleet_seq_mld.py:211-219 */
            for (int k = 0; k <= N; ++k) {
                aff fp = par(P, 0, 0, k), fv = par(P, 0, 1, k), bp = par(P, 1, 0, k), bv = par(P, 1, 1, k);
                aff lp = par(P, 2, 0, k), lv = par(P, 2, 1, k);
                if (!front && !leader) add_track(Q, &Q->p[0][k], &Q->v[0][k], &fp, &fv);
                if (!trailer && !leader) add_track(Q, &bp, &bv, &Q->p[0][k], &Q->v[0][k]);
                if (leader) {
                    if (!real_ref) add_pair(Q, &Q->p[0][k], &Q->v[0][k], &lp, &lv);
                    else add_track(Q, &Q->p[0][k], &Q->v[0][k], &lp, &lv);
                }
                if (!front) bad |= add_safe(Q, &Q->p[0][k], &fp);
                if (!trailer) bad |= add_safe(Q, &bp, &Q->p[0][k]);
                if (leader && real_ref && front) bad |= add_safe(Q, &Q->p[0][k], &lp);
            }
        } break;
        case K_EVENT: {                    /* fleet_event_based.py:143-291 */
            int nf = P->n_front, nb = P->n_behind, me = nf > 0 ? 1 : 0, f1 = me - 1, b1 = me + 1;
            int rl = P->leader_index;
            for (int k = 0; k <= N; ++k) {
                aff lp = par(P, 0, 0, k), lv = par(P, 0, 1, k), f2p = par(P, 1, 0, k), f2v = par(P, 1, 1, k);
                aff b2p = par(P, 2, 0, k), b2v = par(P, 2, 1, k);
                if (rl != NO_LEADER) {
                    int who = rl == -1 ? b1 : (rl == 0 ? me : f1);
                    add_pair(Q, &Q->p[who][k], &Q->v[who][k], &lp, &lv);
                }
                if (nf > 0) { add_track(Q, &Q->p[me][k], &Q->v[me][k], &Q->p[f1][k], &Q->v[f1][k]); bad |= add_safe(Q, &Q->p[me][k], &Q->p[f1][k]); }
                if (nf > 1) { add_track(Q, &Q->p[f1][k], &Q->v[f1][k], &f2p, &f2v); bad |= add_safe(Q, &Q->p[f1][k], &f2p); }
                if (nb > 0) { add_track(Q, &Q->p[b1][k], &Q->v[b1][k], &Q->p[me][k], &Q->v[me][k]); bad |= add_safe(Q, &Q->p[b1][k], &Q->p[me][k]); }
                if (nb > 1) { add_track(Q, &b2p, &b2v, &Q->p[b1][k], &Q->v[b1][k]); bad |= add_safe(Q, &b2p, &Q->p[b1][k]); }
            }
        } break;
        case K_ADMM: {                     /* fleet_naive_admm.py:124-237; extras x_front, x_back */
            int ef = X, eb = X + (front ? 0 : 2 * np1);
            for (int k = 0; k <= N; ++k) {
                aff lp = par(P, 0, 0, k), lv = par(P, 0, 1, k);
                if (!front) {
                    aff cp = aff_var(ef + k), cv = aff_var(ef + np1 + k);
                    if (!leader) add_track(Q, &Q->p[0][k], &Q->v[0][k], &cp, &cv);
                    for (int row = 0; row < 2; ++row) {
                        aff dlt = row == 0 ? cp : cv;
                        dlt.c -= par(P, 2, row, k).c;
                        add_lin(Q, par(P, 1, row, k).c, &dlt);
                        add_res(Q, &dlt, 0.5 * P->rho);
                    }
                    bad |= add_safe(Q, &Q->p[0][k], &cp);
                }
                if (!trailer) {
                    aff cp = aff_var(eb + k), cv = aff_var(eb + np1 + k);
                    if (!leader) add_track(Q, &cp, &cv, &Q->p[0][k], &Q->v[0][k]);
                    for (int row = 0; row < 2; ++row) {
                        aff dlt = row == 0 ? cp : cv;
                        dlt.c -= par(P, 4, row, k).c;
                        add_lin(Q, par(P, 3, row, k).c, &dlt);
                        add_res(Q, &dlt, 0.5 * P->rho);
                    }
                    bad |= add_safe(Q, &cp, &Q->p[0][k]);
                }
                if (leader) add_pair(Q, &Q->p[0][k], &Q->v[0][k], &lp, &lv);
            }
        } break;
        case K_GADMM: {                    /* fleet_g_admm.py:98-158 + consensus terms on the augmented state */
            int nf = P->n_front, nb = P->n_behind, na = nf + nb + 1;
            const double* y = P->params + 2 * np1;
            const double* z = y + (size_t)na * 2 * np1;
            for (int k = 0; k <= N; ++k) {
                aff lp = par(P, 0, 0, k), lv = par(P, 0, 1, k);
                if (leader) add_pair(Q, &Q->p[0][k], &Q->v[0][k], &lp, &lv);
                else {
                    aff cp = aff_var(X + k), cv = aff_var(X + np1 + k);
                    add_track(Q, &Q->p[0][k], &Q->v[0][k], &cp, &cv);
                    bad |= add_safe(Q, &Q->p[0][k], &cp);
                }
                for (int a = 0; a < na; ++a)
                    for (int row = 0; row < 2; ++row) {
                        aff xa;
                        if (a == nf) xa = row == 0 ? Q->p[0][k] : Q->v[0][k];
                        else xa = aff_var(X + (a < nf ? a : a - 1) * 2 * np1 + row * np1 + k);
                        xa.c -= z[(2 * a + row) * np1 + k];
                        add_lin(Q, y[(2 * a + row) * np1 + k], &xa);
                        add_res(Q, &xa, 0.5 * P->rho);
                    }
            }
        } break;
    }
    return bad;
}

typedef struct {
    const mpc_prob* P;
    int use_bound;
    int modes[MAXVEH * MAXN];
    double rlo[MAXVEH][MAXN + 1], rhi[MAXVEH][MAXN + 1];
    double best, second;
    int best_modes[MAXVEH * MAXN];
    double best_z[MAXVAR];
    long leaves, nodes;
    qp_build Q;
    int numeric;
} search;

static int solve_modes(search* S, const int* modes, double* z, double* obj) {
    const mpc_prob* P = S->P;
    /* the k = 0 region rows have zero normals: build_qp resolves them as constant rows */
    if (build_qp(P, modes, &S->Q)) { *obj = INFINITY; return 1; }
    int it = 0;
    int st = hvo_qp_solve(S->Q.n, S->Q.m, S->Q.H, S->Q.g, S->Q.c0, S->Q.A, S->Q.b, S->Q.w, z, NULL, obj, &it);
    if (st >= 2) S->numeric = 1;
    return st;
}

static double g_margin_rel = 1e-6;
void hvo_set_margin(double m) { g_margin_rel = m; }

static void search_rec(search* S, int d) {
    const mpc_prob* P = S->P;
    const int nl = P->nl, N = P->N, D = nl * N;
    double z[MAXVAR], obj;
    double vrel = NAN;
    if (d == D) {
        int st = solve_modes(S, S->modes, z, &obj);
        S->leaves++;
        if (st != 0) return;
        if (obj < S->best) {
            S->second = S->best; S->best = obj;
            memcpy(S->best_modes, S->modes, sizeof(int) * D);
            memcpy(S->best_z, z, sizeof(double) * P->nvar);
        } else if (obj < S->second) S->second = obj;
        return;
    }
    const int i = d % nl, k = d / nl;
    if (S->use_bound && d > 0) {
        int md[MAXVEH * MAXN];
        for (int e = 0; e < D; ++e) {
            int ei = e / N, ek = e % N;          /* modes[] is vehicle-major, decisions stage-major */
            md[e] = (ek * nl + ei < d) ? S->modes[e] : -1;
        }
        int st = solve_modes(S, md, z, &obj);
        S->nodes++;
        if (st == 1) return;
        if (st == 0) {
            double margin = g_margin_rel * fmax(1.0, fabs(S->best));
            if (isfinite(S->best) && obj > S->best + margin) return;
            /* relaxed velocity of the branching stage, to order the children */
            const aff* e = &S->Q.v[i][k];
            vrel = e->c;
            for (int j = 0; j < S->Q.n; ++j) vrel += e->g[j] * z[j];
        }
    }
    int order[MAXMODES], cnt = 0;
    double dist[MAXMODES];
    for (int r = 0; r < P->R; ++r) {
        if (!(P->lo[r] <= P->hi[r])) continue;
        double ds = 0;
        if (vrel == vrel) ds = vrel < P->lo[r] ? P->lo[r] - vrel : (vrel > P->hi[r] ? vrel - P->hi[r] : 0);
        int pos = cnt++;
        while (pos > 0 && dist[pos - 1] > ds) { dist[pos] = dist[pos - 1]; order[pos] = order[pos - 1]; --pos; }
        dist[pos] = ds; order[pos] = r;
    }
    for (int c = 0; c < cnt; ++c) {
        int r = order[c];
        /* reachability pruning: provably empty intersections only */
        double jlo = fmax(S->rlo[i][k], P->lo[r]), jhi = fmin(S->rhi[i][k], P->hi[r]);
        if (jlo > jhi + 1e-9) continue;
        double nlo = fmax(P->a[i][r] * jlo + P->c[i][r] + P->b[i][r] * U_MIN, jlo + A_DEC + k * P->tight);
        double nhi = fmin(P->a[i][r] * jhi + P->c[i][r] + P->b[i][r] * U_MAX, jhi + A_ACC - k * P->tight);
        nlo = fmax(nlo, V_MIN); nhi = fmin(nhi, V_MAX);
        if (nlo > nhi + 1e-9) continue;
        S->rlo[i][k + 1] = nlo - 1e-9; S->rhi[i][k + 1] = nhi + 1e-9;
        S->modes[i * N + k] = r;
        search_rec(S, d + 1);
    }
}

static int prob_init(mpc_prob* P, int kind, int model, int nl, int N, int flags, int leader_index,
                     int n_front, int n_behind, double d0, double t0, double tight, double rho,
                     const double* x0, const double* mass, const double* params) {
    memset(P, 0, sizeof *P);
    P->kind = kind; P->model = model; P->nl = nl; P->N = N; P->flags = flags; P->leader_index = leader_index;
    P->n_front = n_front; P->n_behind = n_behind; P->d0 = d0; P->t0 = t0; P->tight = tight; P->rho = rho;
    P->x0 = x0; P->mass = mass; P->params = params;
    int np1 = N + 1;
    P->ne = 0;
    if (kind == K_ADMM) P->ne = ((flags & F_FRONT) ? 0 : 2 * np1) + ((flags & F_TRAILER) ? 0 : 2 * np1);
    if (kind == K_GADMM) P->ne = (n_front + n_behind) * 2 * np1;
    P->nvar = nl * N + P->ne;
    if (P->nvar > MAXVAR || nl > MAXVEH || N > MAXN) return 1;
    for (int i = 0; i < nl; ++i)
        P->R = hvo_mode_table(model, mass[i], P->a[i], P->b[i], P->c[i], P->lo, P->hi, P->gear);
    return 0;
}

/* Global optimum of one MPC problem.  method: 0 exhaustive (reachability-pruned) enumeration,
 * 1 branch and bound.  fixed_modes (nl*N, vehicle-major) != NULL: the single fixed-mode QP.
 * status: 2 optimal, 3 infeasible, 12 numerical trouble in a QP. */
int hvo_mpc_solve(int kind, int model, int nl, int N, int flags, int leader_index, int n_front,
                  int n_behind, double d0, double t0, double tight, double rho, const double* x0,
                  const double* mass, const double* params, const int32_t* fixed_modes, int method,
                  double* u, double* xtraj, double* extra, int32_t* modes, double* obj,
                  double* second_best, int64_t* leaves, int64_t* nodes) {
    mpc_prob P;
    if (prob_init(&P, kind, model, nl, N, flags, leader_index, n_front, n_behind, d0, t0, tight, rho, x0, mass, params)) return -1;
    search* S = (search*)calloc(1, sizeof(search));
    int n = P.nvar;
    S->P = &P; S->use_bound = method; S->best = INFINITY; S->second = INFINITY;
    S->Q.H = (double*)malloc(sizeof(double) * n * n); S->Q.g = (double*)malloc(sizeof(double) * n);
    S->Q.A = (double*)malloc(sizeof(double) * MAXROWS * n); S->Q.b = (double*)malloc(sizeof(double) * MAXROWS);
    S->Q.w = (double*)malloc(sizeof(double) * MAXROWS);
    for (int i = 0; i < nl; ++i) { S->rlo[i][0] = x0[2 * i + 1]; S->rhi[i][0] = x0[2 * i + 1]; }
    if (fixed_modes) {
        double z[MAXVAR], o;
        int md[MAXVEH * MAXN];
        for (int e = 0; e < nl * N; ++e) md[e] = fixed_modes[e];
        int st = solve_modes(S, md, z, &o);
        S->leaves = 1;
        if (st == 0) { S->best = o; memcpy(S->best_modes, md, sizeof(int) * nl * N); memcpy(S->best_z, z, sizeof(double) * n); }
    } else {
        search_rec(S, 0);
    }
    if (leaves) *leaves = S->leaves;
    if (nodes) *nodes = S->nodes;
    if (second_best) *second_best = S->second;
    int status;
    if (isfinite(S->best)) {
        status = 2;
        *obj = S->best;
        /* rebuild the affine maps of the optimal sequence to report the trajectory */
        build_qp(&P, S->best_modes, &S->Q);
        for (int i = 0; i < nl; ++i)
            for (int k = 0; k <= N; ++k) {
                double p = S->Q.p[i][k].c, v = S->Q.v[i][k].c;
                for (int j = 0; j < n; ++j) { p += S->Q.p[i][k].g[j] * S->best_z[j]; v += S->Q.v[i][k].g[j] * S->best_z[j]; }
                xtraj[(size_t)i * 2 * (N + 1) + k] = p; xtraj[(size_t)i * 2 * (N + 1) + (N + 1) + k] = v;
            }
        for (int e = 0; e < nl * N; ++e) { u[e] = S->best_z[e]; modes[e] = S->best_modes[e]; }
        for (int e = 0; e < P.ne; ++e) if (extra) extra[e] = S->best_z[nl * N + e];
    } else {
        status = S->numeric ? 12 : 3;
        *obj = INFINITY;
        for (int e = 0; e < nl * N; ++e) { u[e] = 0; modes[e] = -1; }
        for (int e = 0; e < nl * 2 * (N + 1); ++e) xtraj[e] = 0;
        for (int e = 0; e < P.ne; ++e) if (extra) extra[e] = 0;
    }
    free(S->Q.H); free(S->Q.g); free(S->Q.A); free(S->Q.b); free(S->Q.w); free(S);
    return status;
}

void hvo_mpc_solve_batch(int batch, int kind, int model, int nl, int N, int flags, int leader_index,
                         int n_front, int n_behind, double d0, double t0, double tight, double rho,
                         int npar, int ne, const double* x0, const double* mass, const double* params,
                         const int32_t* fixed_modes, int method, double* u, double* xtraj, double* extra,
                         int32_t* modes, double* obj, double* second_best, int32_t* status, int64_t* leaves,
                         int64_t* nodes) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < batch; ++i) {
        int64_t lv = 0, nd = 0;
        double sb = 0;
        status[i] = hvo_mpc_solve(kind, model, nl, N, flags, leader_index, n_front, n_behind, d0, t0, tight, rho,
                                  x0 + (size_t)i * nl * 2, mass + (size_t)i * nl, params + (size_t)i * npar,
                                  fixed_modes ? fixed_modes + (size_t)i * nl * N : NULL, method,
                                  u + (size_t)i * nl * N, xtraj + (size_t)i * nl * 2 * (N + 1),
                                  extra ? extra + (size_t)i * ne : NULL, modes + (size_t)i * nl * N, obj + i, &sb,
                                  &lv, &nd);
        if (second_best) second_best[i] = sb;
        if (leaves) leaves[i] = lv;
        if (nodes) nodes[i] = nd;
    }
}

/* dense fixed-mode QP for certification in tests (KKT / HiGHS).  Returns rows, or -1 if trivially infeasible */
int hvo_mpc_build_qp_py(int kind, int model, int nl, int N, int flags, int leader_index, int n_front,
                        int n_behind, double d0, double t0, double tight, double rho, const double* x0,
                        const double* mass, const double* params, const int32_t* modes, double* H, double* g,
                        double* c0, double* A, double* b, double* w) {
    mpc_prob P;
    if (prob_init(&P, kind, model, nl, N, flags, leader_index, n_front, n_behind, d0, t0, tight, rho, x0, mass, params)) return -2;
    qp_build* Q = (qp_build*)calloc(1, sizeof(qp_build));
    Q->H = H; Q->g = g; Q->A = A; Q->b = b; Q->w = w;
    int md[MAXVEH * MAXN];
    for (int e = 0; e < nl * N; ++e) md[e] = modes[e];
    int bad = build_qp(&P, md, Q);
    int m = Q->m;
    *c0 = Q->c0;
    free(Q);
    return bad ? -1 : m;
}
