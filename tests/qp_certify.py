"""Independent certification helpers for convex QPs with hard and L1-soft rows (test code).

    min 1/2 x'Hx + g'x + c0 + sum_{soft i} w_i max(0, a_i'x - b_i)   s.t.  a_i'x <= b_i (hard)

* kkt_residuals: optimality CERTIFICATE -- for a convex QP, (x, lam) satisfying these
  conditions to tolerance proves x optimal irrespective of the algorithm that produced it.
* highs_qp: second opinion from HiGHS (scipy's bundled solver, private binding)
  on the explicit-slack formulation  min ... + w's,  A x - s <= b, s >= 0.
"""
import numpy as np


def kkt_residuals(H, g, A, b, wmax, x, lam):
    hard = ~np.isfinite(wmax)
    r = A @ x - b
    stat = np.abs(H @ x + g + A.T @ lam).max()
    prim = max(0.0, r[hard].max()) if hard.any() else 0.0
    dual = max(0.0, (-lam).max(), (lam[~hard] - wmax[~hard]).max() if (~hard).any() else 0.0)
    # complementarity: hard rows lam*r = 0; soft rows lam in w*subdiff(max(0,r))
    comp = np.abs(lam[hard] * r[hard]).max() if hard.any() else 0.0
    if (~hard).any():
        rs, ls, ws = r[~hard], lam[~hard], wmax[~hard]
        comp = max(comp, np.abs(np.where(rs < 0, ls * rs, (ws - ls) * rs)).max())
    scale = max(1.0, np.abs(g).max())
    return dict(stat=stat / scale, prim=prim, dual=dual / max(1.0, np.abs(lam).max()),
                comp=comp / scale)


def objective(H, g, c0, A, b, wmax, x):
    soft = np.isfinite(wmax)
    r = A @ x - b
    return 0.5 * x @ H @ x + g @ x + c0 + (wmax[soft] * np.maximum(0.0, r[soft])).sum()


def highs_qp(H, g, c0, A, b, wmax):
    """Returns (status_ok, x, obj) from HiGHS on the explicit-slack QP."""
    from scipy.optimize._highspy import _core as hc
    import scipy.sparse as sp

    n = H.shape[0]
    soft = np.where(np.isfinite(wmax))[0]
    ns = len(soft)
    m = A.shape[0]
    # variables [x (free), s >= 0]
    Afull = np.zeros((m, n + ns))
    Afull[:, :n] = A
    for k, i in enumerate(soft):
        Afull[i, n + k] = -1.0
    cost = np.concatenate([g, wmax[soft]])
    Hf = np.zeros((n + ns, n + ns))
    Hf[:n, :n] = H
    h = hc._Highs()
    h.setOptionValue("output_flag", False)
    h.setOptionValue("primal_feasibility_tolerance", 1e-9)
    h.setOptionValue("dual_feasibility_tolerance", 1e-9)
    lp = hc.HighsLp()
    lp.num_col_ = n + ns
    lp.num_row_ = m
    lp.col_cost_ = cost
    lp.col_lower_ = np.concatenate([np.full(n, -hc.kHighsInf), np.zeros(ns)])
    lp.col_upper_ = np.full(n + ns, hc.kHighsInf)
    lp.row_lower_ = np.full(m, -hc.kHighsInf)
    lp.row_upper_ = b
    lp.offset_ = c0
    As = sp.csc_matrix(Afull)
    lp.a_matrix_.format_ = hc.MatrixFormat.kColwise
    lp.a_matrix_.start_ = As.indptr
    lp.a_matrix_.index_ = As.indices
    lp.a_matrix_.value_ = As.data
    h.passModel(lp)
    Hl = sp.csc_matrix(np.tril(Hf))
    hess = hc.HighsHessian()
    hess.dim_ = n + ns
    hess.format_ = hc.HessianFormat.kTriangular
    hess.start_ = Hl.indptr
    hess.index_ = Hl.indices
    hess.value_ = Hl.data
    h.passHessian(hess)
    h.run()
    ok = h.getModelStatus() == hc.HighsModelStatus.kOptimal
    sol = h.getSolution()
    x = np.array(sol.col_value)[:n]
    return ok, x, h.getInfo().objective_function_value
