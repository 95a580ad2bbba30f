"""Seeded synthetic inputs for the local-MIQP bench workload and parity tests.

Scenario distribution follows SURVEY.md 8(d) C2: the reference env's reset distribution
(v in 5..35 m/s, gaps 60-160 m; env.py:83-94) with neighbour predictions extrapolated at
constant velocity (fleet_decent_mld.py:421-428) and a constant-velocity leader reference.
`stress=True` widens it (tight gaps -> active safe-distance slacks, velocities near region
edges) to exercise the soft rows and the branch-and-bound.
"""
import numpy as np

FRONT, LEADER, TRAILER = 1, 2, 4
EDGES = np.array([9.235, 12.855, 16.93, 22.92, 23.315, 32.47])


def platoon_local_problems(rng, n_scen, n, N, leader_index=0, stress=False, hetero=False):
    """Returns dict of arrays for n_scen*n local problems (vehicle-major within scenario)."""
    B = n_scen * n
    v = rng.uniform(5.0, 35.0, (n_scen, n))
    if stress:
        near = rng.random((n_scen, n)) < 0.3
        v = np.where(near, rng.choice(EDGES, (n_scen, n)) + rng.normal(0, 0.3, (n_scen, n)), v)
        v = np.clip(v, 4.2, 45.0)
        gaps = rng.uniform(8.0, 120.0, (n_scen, n))
    else:
        gaps = rng.uniform(60.0, 160.0, (n_scen, n))
    p = 3000.0 - np.cumsum(gaps, axis=1) + gaps[:, :1]
    k = np.arange(N + 1)
    pred = np.stack([p[..., None] + v[..., None] * k, np.repeat(v[..., None], N + 1, -1)], axis=2)
    # pred: (n_scen, n, 2, N+1) constant-velocity extrapolation of every vehicle
    xf = np.zeros((n_scen, n, 2, N + 1)); xb = np.zeros((n_scen, n, 2, N + 1))
    xf[:, 1:] = pred[:, :-1]
    xb[:, :-1] = pred[:, 1:]
    lv = rng.uniform(10.0, 30.0, n_scen)
    lp = p[:, leader_index] + rng.uniform(-40.0, 40.0, n_scen)
    xl = np.zeros((n_scen, n, 2, N + 1))
    xl[:, leader_index, 0] = lp[:, None] + lv[:, None] * k
    xl[:, leader_index, 1] = lv[:, None]
    flags = np.zeros((n_scen, n), dtype=np.int32)
    flags[:, 0] |= FRONT
    flags[:, -1] |= TRAILER
    flags[:, leader_index] |= LEADER
    mass = rng.uniform(700.0, 1000.0, (n_scen, n)) if hetero else np.full((n_scen, n), 800.0)
    x0 = np.stack([p, v], axis=-1)
    return dict(flags=flags.reshape(B), mass=mass.reshape(B), x0=x0.reshape(B, 2),
                xf=xf.reshape(B, 2, N + 1), xb=xb.reshape(B, 2, N + 1), xl=xl.reshape(B, 2, N + 1))
