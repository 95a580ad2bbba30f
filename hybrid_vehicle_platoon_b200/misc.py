"""Scenario configuration: controller parameters, simulation presets, leader trajectories and
spacing policies -- same names, arguments and values as the reference's misc/ package
(misc/common_controller_params.py:14-76, misc/leader_trajectory.py:4-97,
misc/spacing_policy.py:4-37) so that fleet scripts can import them from either place."""
from __future__ import annotations

import numpy as np


# ---- spacing policies (misc/spacing_policy.py) ----------------------------------------------
class SpacingPolicy:
    def spacing(self, x):
        raise NotImplementedError


class ConstantSpacingPolicy(SpacingPolicy):
    """sigma(x) = [-d0, 0]'."""

    def __init__(self, d0: float) -> None:
        self.d = np.array([[-d0, 0]]).T

    def spacing(self, x):
        return self.d


class ConstantTimePolicy(SpacingPolicy):
    """sigma(x) = A x + b = [-t0 v - d0, 0]' (constant time headway)."""

    def __init__(self, d0: float, t0: float) -> None:
        self.A = np.array([[0, -t0], [0, 0]])
        self.b = np.array([[-d0], [0]])

    def spacing(self, x):
        x = np.asarray(x)
        if x.ndim == 1:
            x = x.reshape(x.shape[0], 1)
        return self.A @ x + self.b


def spacing_params(policy) -> tuple[float, float]:
    """(d0, t0) of a spacing policy object -- ours or the reference's (duck-typed)."""
    if hasattr(policy, "A") and hasattr(policy, "b"):
        return float(-np.asarray(policy.b)[0, 0]), float(-np.asarray(policy.A)[0, 1])
    if hasattr(policy, "d"):
        return float(-np.asarray(policy.d)[0, 0]), 0.0
    raise TypeError(f"unsupported spacing policy {type(policy).__name__}")


# ---- leader trajectories (misc/leader_trajectory.py) ----------------------------------------
class LeaderTrajectory:
    def __init__(self, trajectory_len: int, ts: float) -> None:
        self.trajectory_len = trajectory_len
        self.ts = ts

    def get_leader_trajectory(self) -> np.ndarray:
        raise NotImplementedError


class ConstantVelocityLeaderTrajectory(LeaderTrajectory):
    def __init__(self, p: float, v: float, trajectory_len: int, ts: float) -> None:
        super().__init__(trajectory_len, ts)
        self.p0, self.v = p, v

    def get_leader_trajectory(self) -> np.ndarray:
        x = np.zeros((2, self.trajectory_len))
        x[0, 0], x[1, 0] = self.p0, self.v
        for k in range(self.trajectory_len - 1):   # sequential sums, as the reference accumulates
            x[0, k + 1] = x[0, k] + self.ts * self.v
            x[1, k + 1] = x[1, k] + self.ts * 0
        return x


class StopAndGoLeaderTrajectory(LeaderTrajectory):
    """vh -> vl at v_change_steps[0] -> vf (or vh) at v_change_steps[1]; the position at the step
    of a change is still advanced with the old velocity (leader_trajectory.py:62-65)."""

    def __init__(self, p, vh, vl, v_change_steps, trajectory_len, ts, vf=None) -> None:
        super().__init__(trajectory_len, ts)
        if len(v_change_steps) != 2:
            raise ValueError(f"v_change_steps should have 2 items, received {len(v_change_steps)}")
        self.p0, self.vh, self.vl, self.vf, self.v_change_steps = p, vh, vl, vf, v_change_steps

    def get_leader_trajectory(self) -> np.ndarray:
        x = np.zeros((2, self.trajectory_len))
        x[0, 0], x[1, 0] = self.p0, self.vh
        v = self.vh
        for k in range(self.trajectory_len - 1):
            x[0, k + 1] = x[0, k] + self.ts * v
            x[1, k + 1] = x[1, k] + self.ts * 0
            if self.v_change_steps[0] <= k < self.v_change_steps[1]:
                v = self.vl
                x[1, k + 1] = v
            elif k >= self.v_change_steps[1]:
                v = self.vh if self.vf is None else self.vf
                x[1, k + 1] = v
        return x


# ---- controller parameters and simulation presets (misc/common_controller_params.py) --------
class Params:
    Q_x = np.diag([1, 0.1])
    Q_u = 1 * np.eye(1)
    q_du = 0
    Q_du = q_du * np.eye(1)
    w = 1e4
    ts = 1
    a_acc = 2.5
    a_dec = -2
    d_safe = 25


class Sim:
    open_loop = False
    real_vehicle_as_reference = False
    vehicle_model_type = "pwa_gear"
    start_from_platoon = False
    quadratic_cost = True
    n = 3
    N = 6
    ep_len = N if open_loop else 150
    spacing_policy = ConstantSpacingPolicy(50)
    leader_trajectory = ConstantVelocityLeaderTrajectory(p=3000, v=20, trajectory_len=ep_len + 50, ts=Params.ts)
    masses = None
    id = f"default_n_{n}_N_{N}"


class Sim_n_task_1(Sim):
    def __init__(self, n: int) -> None:
        self.n = n
        self.id = f"task_1_n_{n}_N_{self.N}"
        self.spacing_policy = ConstantSpacingPolicy(50)
        self.leader_trajectory = ConstantVelocityLeaderTrajectory(
            p=3100, v=20, trajectory_len=self.ep_len + 50, ts=Params.ts)


class Sim_n_task_2(Sim):
    def __init__(self, n: int, seed: int, leader_index: int | None = None, N: int = 6) -> None:
        self.n = n
        self.N = N
        self.id = f"task_2_n_{n}_N_{self.N}" if Params.q_du == 0 else f"task_2_n_{n}_N_{self.N}_q_{Params.q_du}"
        if leader_index is not None:
            self.id += f"_lead_{leader_index}"
        self.spacing_policy = ConstantTimePolicy(10, 3)
        self.leader_trajectory = StopAndGoLeaderTrajectory(
            p=3000, vh=20, vl=10, vf=30, v_change_steps=[30, 50], trajectory_len=self.ep_len + 50, ts=Params.ts)
        np.random.seed(seed)                                  # legacy global RNG, as the reference (Q2)
        self.masses = np.random.uniform(700, 1000, n).tolist()
