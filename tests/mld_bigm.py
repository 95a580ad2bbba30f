"""INDEPENDENT checker for the MIQP half of the hot path: the reference's MLD model written out explicitly
(binaries delta, auxiliaries z, big-M rows) and handed to somebody else's solver.  Test code only.

Every other MIQP check in this repo (oracle/hvp_oracle*.c, the CUDA kernels, the CPU port) shares one restatement:
"the MIQP optimum is the minimum over PWA mode sequences of a fixed-mode QP in velocity / input space".  This module
does NOT: it builds the model the way dmpcpwa's MpcMld does (index ranges pinned by the in-tree
mirror mpcs/mpc_gear.py:207-235) --

    x (nx, N+1), u (nu, N) free;  z (s, nx, N) free;  delta (s, N) binary;  sum_r delta[r, k] = 1
    D x_k <= E  (k = 1..N; constrain_first_state=False),  F u_k <= G  (k < N),  x_{k+1} = sum_r z[r, :, k]
    S_r x_k + R_r u_k - T_r <= Mstar_r (1 - delta[r, k])
    z[r,:,k] <= M_ub delta,  z >= M_lb delta,  z <= A_r x_k + B_r u_k + c_r - M_lb (1 - delta),
    z >= A_r x_k + B_r u_k + c_r - M_ub (1 - delta),     x_0 = state

with the big-M constants from the box {D x <= E, F u <= G}, on the system dicts {S,R,T,A,B,c,D,E,F,G} of
models.py:397-492 -- and adds the controller's cost and rows term by term from the reference files:
LocalMpcMld fleet_decent_mld.py:61-208, MpcMldCent mpcs/cent_mld.py:48-177.

Solvers: `scipy.optimize.milp` (HiGHS branch and bound) for the 1-norm cost (quadratic_cost=False,
cent_mld.py:58-61 -> sum |Q e| through epigraph variables); for the 2-norm cost HiGHS refuses MIQPs, so the binaries
are fixed to each of the s^N mode sequences in turn and HiGHS solves the remaining convex QP of the SAME big-M model
(exhaustive: small N only).  Either way the big-M side effects on inactive regions that DESIGN.md 7 argues away are
part of the model here."""
import itertools

import numpy as np
import scipy.sparse as sp
from scipy.optimize import Bounds, LinearConstraint, milp

QX = np.array([1.0, 0.1])       # Params.Q_x diagonal (misc/common_controller_params.py:15)
QU, W_SLACK, A_ACC, A_DEC, D_SAFE = 1.0, 1e4, 2.5, -2.0, 25.0


class _Model:
    """Tiny LP/QP modelling layer: named variable blocks, <= / == rows, linear + diagonal-free quadratic cost."""

    def __init__(self):
        self.n = 0
        self.lb, self.ub, self.integ = [], [], []
        self.rows, self.lo, self.hi = [], [], []          # each row: dict {var: coef}
        self.c = {}
        self.c0 = 0.0
        self.quad = []                                     # list of (row dict, const, weight): weight * (row'x + const)^2

    def var(self, shape, lb=-np.inf, ub=np.inf, integer=False):
        size = int(np.prod(shape))
        idx = np.arange(self.n, self.n + size).reshape(shape)
        self.n += size
        self.lb += [lb] * size; self.ub += [ub] * size; self.integ += [int(integer)] * size
        return idx

    def add(self, row, lo, hi):
        self.rows.append(row); self.lo.append(lo); self.hi.append(hi)

    def le(self, row, rhs):
        self.add(row, -np.inf, rhs)

    def eq(self, row, rhs):
        self.add(row, rhs, rhs)

    def cost(self, row, const=0.0):
        for k, v in row.items():
            self.c[k] = self.c.get(k, 0.0) + v
        self.c0 += const

    def abs_cost(self, row, const, weight):
        """weight * |row'x + const| through an epigraph variable (dmpcpwa min_1_norm)."""
        if weight == 0.0:
            return
        t = int(self.var((1,), lb=0.0)[0])
        r1 = dict(row); r1[t] = r1.get(t, 0.0) - 1.0
        self.le(r1, -const)                                # row + const <= t
        r2 = {k: -v for k, v in row.items()}; r2[t] = r2.get(t, 0.0) - 1.0
        self.le(r2, const)                                 # -(row + const) <= t
        self.cost({t: weight})

    def sq_cost(self, row, const, weight):
        if weight != 0.0:
            self.quad.append((row, const, weight))

    def matrices(self):
        ri, ci, vv = [], [], []
        for i, r in enumerate(self.rows):
            for k, v in r.items():
                ri.append(i); ci.append(k); vv.append(v)
        A = sp.csr_matrix((vv, (ri, ci)), shape=(len(self.rows), self.n))
        c = np.zeros(self.n)
        for k, v in self.c.items():
            c[k] = v
        return A, np.array(self.lo), np.array(self.hi), c

    def hessian(self):
        """(H, g, c0) of the quadratic terms: sum w (r'x + k)^2 = x'(w r r')x + 2 w k r'x + w k^2."""
        H = np.zeros((self.n, self.n)); g = np.zeros(self.n); c0 = 0.0
        for row, k, wt in self.quad:
            idx = np.array(list(row.keys())); val = np.array(list(row.values()))
            H[np.ix_(idx, idx)] += 2.0 * wt * np.outer(val, val)
            g[idx] += 2.0 * wt * k * val
            c0 += wt * k * k
        return H, g, c0


def _box_max(a, lo, hi):
    """max of a'y over the box lo <= y <= hi."""
    return float(np.where(a > 0, a * hi, a * lo).sum())


def _add_mld(M, sysd, N, x0):
    """One PWA subsystem in MLD form (docstring above).  Returns index arrays (x, u, delta)."""
    S, R, T, A, B, c = (sysd[k] for k in "SRTABc")
    D, E, F, G = (np.asarray(sysd[k], dtype=np.float64) for k in "DEFG")
    s, nx, nu = len(A), 2, 1
    x = M.var((nx, N + 1)); u = M.var((nu, N)); z = M.var((s, nx, N)); dl = M.var((s, N), 0.0, 1.0, integer=True)
    # the box of (x, u) the big-M constants are taken over: D x <= E, F u <= G (rows are +-e_i here)
    xlo = np.full(nx, -np.inf); xhi = np.full(nx, np.inf)
    for row, e in zip(D, E.ravel()):
        i = int(np.argmax(np.abs(row)))
        if row[i] > 0: xhi[i] = min(xhi[i], e / row[i])
        else: xlo[i] = max(xlo[i], e / row[i])
    ulo = np.full(nu, -np.inf); uhi = np.full(nu, np.inf)
    for row, g in zip(F, G.ravel()):
        i = int(np.argmax(np.abs(row)))
        if row[i] > 0: uhi[i] = min(uhi[i], g / row[i])
        else: ulo[i] = max(ulo[i], g / row[i])
    ylo, yhi = np.r_[xlo, ulo], np.r_[xhi, uhi]
    Mub = np.full(nx, -np.inf); Mlb = np.full(nx, np.inf)
    for r in range(s):
        AB = np.hstack([np.asarray(A[r], float), np.asarray(B[r], float)])
        for i in range(nx):
            Mub[i] = max(Mub[i], _box_max(AB[i], ylo, yhi) + float(np.asarray(c[r]).ravel()[i]))
            Mlb[i] = min(Mlb[i], -_box_max(-AB[i], ylo, yhi) + float(np.asarray(c[r]).ravel()[i]))
    for k in range(N):
        M.eq({int(dl[r, k]): 1.0 for r in range(s)}, 1.0)
        for i, (row, g) in enumerate(zip(F, G.ravel())):
            M.le({int(u[j, k]): row[j] for j in range(nu) if row[j] != 0.0}, g)
        for i in range(nx):
            row = {int(x[i, k + 1]): 1.0}
            for r in range(s):
                row[int(z[r, i, k])] = -1.0
            M.eq(row, 0.0)
        for r in range(s):
            Sr, Rr, Tr = np.asarray(S[r], float), np.asarray(R[r], float), np.asarray(T[r], float).ravel()
            for i in range(Sr.shape[0]):
                ms = _box_max(np.r_[Sr[i], Rr[i]], ylo, yhi) - Tr[i]
                row = {int(x[j, k]): Sr[i, j] for j in range(nx) if Sr[i, j] != 0.0}
                row.update({int(u[j, k]): Rr[i, j] for j in range(nu) if Rr[i, j] != 0.0})
                row[int(dl[r, k])] = row.get(int(dl[r, k]), 0.0) + ms
                M.le(row, Tr[i] + ms)                                   # S x + R u - T <= M* (1 - delta)
            Ar, Br, cr = np.asarray(A[r], float), np.asarray(B[r], float), np.asarray(c[r], float).ravel()
            for i in range(nx):
                zi, di = int(z[r, i, k]), int(dl[r, k])
                M.le({zi: 1.0, di: -Mub[i]}, 0.0)                       # z <= M_ub delta
                M.le({zi: -1.0, di: Mlb[i]}, 0.0)                       # z >= M_lb delta
                aff = {int(x[j, k]): Ar[i, j] for j in range(nx) if Ar[i, j] != 0.0}
                aff.update({int(u[j, k]): Br[i, j] for j in range(nu) if Br[i, j] != 0.0})
                r1 = {zi: 1.0, di: -Mlb[i]}
                for kk, v in aff.items(): r1[kk] = r1.get(kk, 0.0) - v
                M.le(r1, cr[i] - Mlb[i])                                # z <= A x + B u + c - M_lb (1 - delta)
                r2 = {zi: -1.0, di: Mub[i]}
                for kk, v in aff.items(): r2[kk] = r2.get(kk, 0.0) + v
                M.le(r2, -cr[i] + Mub[i])                               # z >= A x + B u + c - M_ub (1 - delta)
    for k in range(1, N + 1):
        for row, e in zip(D, E.ravel()):
            M.le({int(x[j, k]): row[j] for j in range(nx) if row[j] != 0.0}, e)
    for i in range(nx):
        M.eq({int(x[i, 0]): 1.0}, float(x0[i]))
    return x, u, dl


def _norm_cost(M, quadratic, rows_consts, weights):
    """cost_func(e, Q) for e given as [(row, const)] per component: e'Qe or sum |Q e| (env.py:118-124 mirrors)."""
    for (row, const), wq in zip(rows_consts, weights):
        if quadratic:
            M.sq_cost(row, const, wq)
        else:
            M.abs_cost(row, const, wq)


def _spacing(d0, t0, xi, k):
    """spacing_policy.spacing(x_k) = [-t0 v - d0, 0] as (row, const) per component (misc/spacing_policy.py:13-37)."""
    return [({int(xi[1, k]): -t0}, -d0), ({}, 0.0)]


def _sub(a, b):
    row = dict(a[0])
    for k, v in b[0].items():
        row[k] = row.get(k, 0.0) - v
    return ({k: v for k, v in row.items() if v != 0.0}, a[1] - b[1])


def build_local(sysd, N, x0, xf, xb, xl, *, is_front, is_leader, is_trailer, d0=50.0, t0=0.0, tight=0.0, quadratic=True):
    """LocalMpcMld (fleet_decent_mld.py:61-208) on the explicit MLD model."""
    M = _Model()
    x, u, dl = _add_mld(M, sysd, N, x0)
    sf = M.var((N + 1,), 0.0, 0.0 if is_front else np.inf)
    sb = M.var((N + 1,), 0.0, 0.0 if is_trailer else np.inf)
    st = lambda i, k: ({int(x[i, k]): 1.0}, 0.0)
    for k in range(N + 1):
        if not is_front and not is_leader:            # x - x_front - spacing(x)
            sp_ = _spacing(d0, t0, x, k)
            e = [_sub(_sub(st(i, k), ({}, float(xf[i, k]))), sp_[i]) for i in range(2)]
            _norm_cost(M, quadratic, e, QX)
        if not is_trailer and not is_leader:          # x_back - x - spacing(x_back)
            spb = [-t0 * float(xb[1, k]) - d0, 0.0]
            e = [_sub(({}, float(xb[i, k]) - spb[i]), st(i, k)) for i in range(2)]
            _norm_cost(M, quadratic, e, QX)
        if is_leader:                                 # x - leader_x
            e = [_sub(st(i, k), ({}, float(xl[i, k]))) for i in range(2)]
            _norm_cost(M, quadratic, e, QX)
        M.cost({int(sf[k]): W_SLACK, int(sb[k]): W_SLACK})
        if not is_front:
            M.le({int(x[0, k]): 1.0, int(sf[k]): -1.0}, float(xf[0, k]) - D_SAFE)
        if not is_trailer:
            M.le({int(x[0, k]): -1.0, int(sb[k]): -1.0}, -float(xb[0, k]) - D_SAFE)
    for k in range(N):
        _norm_cost(M, quadratic, [({int(u[0, k]): 1.0}, 0.0)], [QU])
        M.le({int(x[1, k + 1]): -1.0, int(x[1, k]): 1.0}, -A_DEC - k * tight)       # a_dec <= dv - k tight
        M.le({int(x[1, k + 1]): 1.0, int(x[1, k]): -1.0}, A_ACC - k * tight)
    return M, x, u, dl


def build_cent(systems, N, x0, leader_traj, *, leader_index=0, d0=50.0, t0=0.0, tight=0.0, quadratic=True):
    """MpcMldCent (mpcs/cent_mld.py:48-177; real_vehicle_as_reference=False) on the explicit MLD model of n
    decoupled subsystems (MpcMldCentDecup)."""
    n = len(systems)
    M = _Model()
    xs, us, ds = [], [], []
    for i in range(n):
        x, u, dl = _add_mld(M, systems[i], N, x0[i])
        xs.append(x); us.append(u); ds.append(dl)
    s = M.var((n, N + 1), 0.0, np.inf)
    st = lambda v, i, k: ({int(xs[v][i, k]): 1.0}, 0.0)
    for k in range(N + 1):
        e = [_sub(st(leader_index, i, k), ({}, float(leader_traj[i, k]))) for i in range(2)]
        _norm_cost(M, quadratic, e, QX)
        for v in range(1, n):
            sp_ = _spacing(d0, t0, xs[v], k)
            e = [_sub(_sub(st(v, i, k), st(v - 1, i, k)), sp_[i]) for i in range(2)]
            _norm_cost(M, quadratic, e, QX)
            M.le({int(xs[v][0, k]): 1.0, int(xs[v - 1][0, k]): -1.0, int(s[v, k]): -1.0}, -D_SAFE)
        for v in range(n):
            M.cost({int(s[v, k]): W_SLACK})
    for v in range(n):
        for k in range(N):
            _norm_cost(M, quadratic, [({int(us[v][0, k]): 1.0}, 0.0)], [QU])
            M.le({int(xs[v][1, k + 1]): -1.0, int(xs[v][1, k]): 1.0}, -A_DEC - k * tight)
            M.le({int(xs[v][1, k + 1]): 1.0, int(xs[v][1, k]): -1.0}, A_ACC - k * tight)
    return M, xs, us, ds


def solve_milp(M, time_limit=60.0):
    """HiGHS branch and bound on a model with linear cost (1-norm variant).  Returns (status_ok, x, obj)."""
    assert not M.quad
    A, lo, hi, c = M.matrices()
    res = milp(c, constraints=LinearConstraint(A, lo, hi), integrality=np.array(M.integ),
               bounds=Bounds(np.array(M.lb), np.array(M.ub)),
               options=dict(mip_rel_gap=0.0, time_limit=time_limit, presolve=True))
    if res.status != 0 or res.x is None:
        return False, None, np.inf
    return True, res.x, float(res.fun) + M.c0


def solve_qp_fixed(M, fixed, tol=1e-9):
    """HiGHS on the convex QP left when the binaries are fixed (`fixed`: {var index: 0/1}).  (ok, x, obj)."""
    from scipy.optimize._highspy import _core as hc
    A, lo, hi, c = M.matrices()
    H, g, c0 = M.hessian()
    lb, ub = np.array(M.lb, dtype=np.float64), np.array(M.ub, dtype=np.float64)
    for k, v in fixed.items():
        lb[k] = ub[k] = float(v)
    h = hc._Highs()
    h.setOptionValue("output_flag", False)
    h.setOptionValue("primal_feasibility_tolerance", tol)
    h.setOptionValue("dual_feasibility_tolerance", tol)
    lp = hc.HighsLp()
    lp.num_col_, lp.num_row_ = M.n, A.shape[0]
    lp.col_cost_ = c + g
    lp.col_lower_ = np.where(np.isfinite(lb), lb, -hc.kHighsInf)
    lp.col_upper_ = np.where(np.isfinite(ub), ub, hc.kHighsInf)
    lp.row_lower_ = np.where(np.isfinite(lo), lo, -hc.kHighsInf)
    lp.row_upper_ = np.where(np.isfinite(hi), hi, hc.kHighsInf)
    lp.offset_ = M.c0 + c0
    As = sp.csc_matrix(A)
    lp.a_matrix_.format_ = hc.MatrixFormat.kColwise
    lp.a_matrix_.start_, lp.a_matrix_.index_, lp.a_matrix_.value_ = As.indptr, As.indices, As.data
    h.passModel(lp)
    Hl = sp.csc_matrix(np.tril(H))
    hess = hc.HighsHessian()
    hess.dim_ = M.n
    hess.format_ = hc.HessianFormat.kTriangular
    hess.start_, hess.index_, hess.value_ = Hl.indptr, Hl.indices, Hl.data
    h.passHessian(hess)
    h.run()
    if h.getModelStatus() != hc.HighsModelStatus.kOptimal:
        return False, None, np.inf
    return True, np.array(h.getSolution().col_value), float(h.getInfo().objective_function_value)


def enumerate_miqp(M, deltas, reach=None):
    """2-norm MIQP by brute force over the mode sequences of ONE or several subsystems: delta arrays (s, N) each.
    `reach(seq_tuple) -> bool` may skip sequences (None: all s^N per subsystem).  Returns (best obj, best modes,
    second-best obj, number of feasible sequences)."""
    per = []
    for dl in deltas:
        s, N = dl.shape
        per.append(list(itertools.product(range(s), repeat=N)))
    best, bm, second, feas = np.inf, None, np.inf, 0
    bx = None
    for combo in itertools.product(*per):
        if reach is not None and not reach(combo):
            continue
        fixed = {}
        for dl, seq in zip(deltas, combo):
            s, N = dl.shape
            for k in range(N):
                for r in range(s):
                    fixed[int(dl[r, k])] = 1.0 if seq[k] == r else 0.0
        ok, x, obj = solve_qp_fixed(M, fixed)
        if not ok:
            continue
        feas += 1
        if obj < best:
            second, best, bm, bx = best, obj, combo, x
        elif obj < second:
            second = obj
    return best, bm, second, feas, bx
