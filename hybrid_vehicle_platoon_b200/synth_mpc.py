"""Seeded synthetic inputs for the compiled-MPC parity tests (cent / event / ADMM / g-ADMM / gears)."""
import numpy as np

FRONT, LEADER, TRAILER = 1, 2, 4
CENT, LOCAL, EVENT, ADMM, GADMM = 1, 2, 3, 4, 5
EDGES = np.array([9.235, 12.855, 16.93, 22.92, 23.315, 32.47])


def platoon_states(rng, B, nl, stress=False, vlo=5.0, vhi=35.0):
    v = rng.uniform(vlo, vhi, (B, nl))
    if stress:
        near = rng.random((B, nl)) < 0.3
        v = np.where(near, rng.choice(EDGES, (B, nl)) + rng.normal(0, 0.3, (B, nl)), v)
        v = np.clip(v, 4.2, 45.0)
        gaps = rng.uniform(8.0, 120.0, (B, nl))
    else:
        gaps = rng.uniform(40.0, 160.0, (B, nl))
    p = 3000.0 - np.cumsum(gaps, axis=1)
    return np.stack([p, v], axis=-1)          # (B, nl, 2)


def const_vel(x, N):
    """(..., 2) -> (..., 2, N+1) constant-velocity extrapolation (fleet_decent_mld.py:421-428)."""
    k = np.arange(N + 1)
    return np.stack([x[..., :1] + x[..., 1:] * k, np.repeat(x[..., 1:], N + 1, -1)], axis=-2)


def cent_cases(rng, B, n, N, stress=False, leader_index=0):
    x0 = platoon_states(rng, B, n, stress)
    lead = x0[:, leader_index].copy()
    lead[:, 0] += rng.uniform(-30, 30, B)
    lead[:, 1] = rng.uniform(10, 30, B)
    params = const_vel(lead, N).reshape(B, -1)
    return x0, params


def event_cases(rng, B, nf, nb, N, stress=False):
    nl = (nf > 0) + 1 + (nb > 0)
    full = platoon_states(rng, B, nl + 2, stress)       # [f2, local..., b2]
    x0 = full[:, 1:1 + nl]
    lead = x0[:, 0].copy(); lead[:, 0] += rng.uniform(-20, 40, B); lead[:, 1] = rng.uniform(10, 30, B)
    params = np.concatenate([const_vel(lead, N).reshape(B, -1), const_vel(full[:, 0], N).reshape(B, -1),
                             const_vel(full[:, -1], N).reshape(B, -1)], axis=1)
    return x0, params


def admm_cases(rng, B, N, stress=False):
    full = platoon_states(rng, B, 3, stress)            # [front, me, back]
    x0 = full[:, 1:2]
    lead = x0[:, 0].copy(); lead[:, 0] += rng.uniform(-20, 40, B); lead[:, 1] = rng.uniform(10, 30, B)
    zf = const_vel(full[:, 0], N) + rng.normal(0, 1.0, (B, 2, N + 1))
    zb = const_vel(full[:, 2], N) + rng.normal(0, 1.0, (B, 2, N + 1))
    yf = rng.normal(0, 5.0, (B, 2, N + 1)); yb = rng.normal(0, 5.0, (B, 2, N + 1))
    params = np.concatenate([a.reshape(B, -1) for a in (const_vel(lead, N), yf, zf, yb, zb)], axis=1)
    return x0, params


def gadmm_cases(rng, B, nf, nb, N, stress=False):
    na = nf + nb + 1
    full = platoon_states(rng, B, na, stress)
    x0 = full[:, nf:nf + 1]
    lead = x0[:, 0].copy(); lead[:, 0] += rng.uniform(-20, 40, B); lead[:, 1] = rng.uniform(10, 30, B)
    z = const_vel(full, N) + rng.normal(0, 1.0, (B, na, 2, N + 1))
    y = rng.normal(0, 5.0, (B, na, 2, N + 1))
    params = np.concatenate([const_vel(lead, N).reshape(B, -1), y.reshape(B, -1), z.reshape(B, -1)], axis=1)
    # a plausible fixed sequence: region of the constant-velocity roll-out
    modes = np.clip(np.digitize(np.repeat(x0[:, :, 1:], N, -1), np.r_[EDGES]), 0, 6).astype(np.int32)
    return x0, params, modes
