"""GPU-backed drop-in for the reference's PlatoonEnv (env.py:12-219): same constructor, reset(),
step(), get_state(), get_previous_state() and attributes; the stage cost, the violation flag, the
gear derivation and the ten-sub-step hybrid rollout all run in the CUDA rollout kernel
(csrc/rollout.cu) through hvp_rollout_step_host.  BatchedPlatoonEnv steps many scenarios at once
on device tensors."""
from __future__ import annotations

from typing import Any

import numpy as np

from . import api
from ._lib import default_context
from .misc import ConstantSpacingPolicy, ConstantVelocityLeaderTrajectory, spacing_params
from .models import Platoon

_ERR_TEXT = {1: "Velocity {v} of vehicle exeeds true model bounds (2.0706, 59.9715).",
             2: "Gear value out of range 1 - 6.",
             3: "Velocity out of range for gear."}


def raise_rollout_error(err: int) -> None:
    """Surface the kernel's per-scenario error code as the exception the reference raises
    (models.py:32-42,119-122)."""
    if err:
        code, veh, sub = err & 0xFF, (err >> 8) & 0xFF, err >> 16
        raise RuntimeError(f"{_ERR_TEXT.get(code, 'rollout error').format(v='?')} "
                           f"[vehicle {veh}, sub-step {sub}, code {code}]")


class PlatoonEnv:
    """An env for a platoon of non-linear hybrid vehicles who track each other (GPU rollout)."""

    Q_x = np.diag([1, 0.1])
    Q_u = 1 * np.eye(1)
    Q_du = 0 * np.eye(1)
    nx_l = Platoon.nx_l
    nu_l = Platoon.nu_l
    step_counter = 0
    viol_counter: list = []   # class attribute shared by all instances, as in the reference (Q7)

    def __init__(self, n: int, platoon, ep_len: int, leader_index: int = 0, ts: float = 1,
                 leader_trajectory=ConstantVelocityLeaderTrajectory(p=3000, v=20, trajectory_len=150, ts=1),
                 spacing_policy=ConstantSpacingPolicy(50), d_safe: float = 25,
                 start_from_platoon: bool = False, quadratic_cost: bool = True,
                 real_vehicle_as_reference: bool = False, ctx=None) -> None:
        if ts != 1:
            raise NotImplementedError("the rollout kernel implements the reference's ts = 1")
        if leader_index != 0 and real_vehicle_as_reference:
            raise NotImplementedError("Not implemented for real vehicle with leader not 0.")
        self.leader_index, self.platoon, self.ts, self.n = leader_index, platoon, ts, n
        self.ep_len, self.d_safe, self.start_from_platoon = ep_len, d_safe, start_from_platoon
        self.leader_trajectory, self.spacing_policy = leader_trajectory, spacing_policy
        self.real_vehicle_as_reference, self.quadratic_cost = real_vehicle_as_reference, quadratic_cost
        self.d0, self.t0 = spacing_params(spacing_policy)
        self.masses = np.array([float(v.m) for v in platoon.vehicles]) if hasattr(platoon, "vehicles") \
            else np.asarray(platoon, dtype=np.float64)
        self.previous_action = None
        self.previous_state = None
        self._ctx = ctx

    # -- reset: bug-for-bug with env.py:70-116 (Q1 int64 state, Q2 legacy global RNG) ---------
    def reset(self, *, seed: int = None, options: dict[str, Any] = None):
        self.leader_x = self.leader_trajectory.get_leader_trajectory()
        self.x = np.tile(np.array([[0], [0]]), (self.n, 1))          # int64, floats truncate (Q1)
        np.random.seed(seed)
        starting_velocities = [30 * np.random.random() + 5 for _ in range(100)]
        starting_positions = [3000.0]
        for _ in range(1, 100):
            starting_positions.append(-100 * np.random.random() + starting_positions[-1] - 60)
        if not self.start_from_platoon:
            for i in range(self.n):
                init_pos = max(starting_positions)
                self.x[i * self.nx_l, :] = init_pos
                self.x[i * self.nx_l + 1, :] = starting_velocities[i]
                starting_positions.remove(init_pos)
        else:
            for i in range(self.n):
                k = (i + 1) if self.real_vehicle_as_reference else i
                self.x[i * self.nx_l:(i + 1) * self.nx_l, :] = \
                    self.leader_x[:, [0]] + k * self.spacing_policy.spacing(self.leader_x[:, [0]])
        self.step_counter = 0
        self.viol_counter.append(np.zeros(self.ep_len))
        return self.x, {}

    def step(self, action: np.ndarray):
        n = self.n
        if action.shape != (n * self.nu_l, 1) and action.shape != (2 * n * self.nu_l, 1):
            raise ValueError(f"Expected action of size {(n * self.nu_l, 1)} (no gears) or "
                             f"{(2 * n * self.nu_l, 1)} (with gears). Got {action.shape}")
        gears = None
        if action.shape[0] == 2 * n:
            u, gears = action[:n, :], action[n:, :].astype(np.int32).reshape(1, n)
        else:
            u = action
            # gears derived from the velocities exist for PWA-gear vehicles only (Platoon.get_gear_from_vehicle_velocity,
            # models.py:259-264 raises for any other vehicle type)
            from .models import PwaGearVehicle
            for i, veh in enumerate(getattr(self.platoon, "vehicles", [])):
                if not isinstance(veh, PwaGearVehicle):
                    raise RuntimeError(f"Gear from velocity asked but the given vehicle {i} is not a PWA vehicle.")
        xo, cost, viol, err = api.rollout_step(
            np.asarray(self.x, dtype=np.float64).reshape(1, 2 * n), u.reshape(1, n), gears, self.masses,
            self.leader_x[:, self.step_counter].reshape(1, 2), d0=self.d0, t0=self.t0,
            leader_index=self.leader_index, d_safe=self.d_safe, quadratic=self.quadratic_cost,
            real_ref=self.real_vehicle_as_reference, ctx=self._ctx)
        if viol[0] and self.step_counter < len(self.viol_counter[-1]):
            self.viol_counter[-1][self.step_counter] = 100
        self.previous_action, self.previous_state = u, self.x
        raise_rollout_error(int(err[0]))
        r = np.array([[cost[0]]])
        self.x = xo.reshape(2 * n, 1)
        self.step_counter += 1
        return self.x, r, False, False, {}

    def get_stage_cost(self, state, action) -> float:
        """env.py:126-180 on the GPU (state/action need not be the env's own), with the reference's side effects: the
        violation counter of the current step and previous_action / previous_state (env.py:165-179)."""
        _, cost, viol, _ = api.rollout_step(
            np.asarray(state, dtype=np.float64).reshape(1, -1), np.asarray(action, dtype=np.float64).reshape(1, -1),
            None, self.masses, self.leader_x[:, self.step_counter].reshape(1, 2), d0=self.d0, t0=self.t0,
            leader_index=self.leader_index, d_safe=self.d_safe, quadratic=self.quadratic_cost,
            real_ref=self.real_vehicle_as_reference, ctx=self._ctx)
        if viol[0] and self.step_counter < len(self.viol_counter[-1]):
            self.viol_counter[-1][self.step_counter] = 100
        self.previous_action, self.previous_state = action, state
        return np.array([[cost[0]]])

    def get_state(self):
        return self.x

    def get_previous_state(self):
        return self.previous_state if self.previous_state is not None else self.x


class BatchedPlatoonEnv:
    """B independent platoons stepped together on the device (scenario-sharded sweeps, C4/C5).
    State is a torch CUDA tensor (B, 2n); step() is asynchronous on torch's current stream."""

    def __init__(self, n, masses=None, leader_index=0, spacing_policy=ConstantSpacingPolicy(50), d_safe=25.0,
                 quadratic_cost=True, device=0, ctx=None):
        import torch
        self.torch = torch
        self.n, self.device = n, torch.device("cuda", device)
        self.ctx = ctx or default_context(device)
        d0, t0 = spacing_params(spacing_policy)
        self.mass = None
        per = False
        if masses is not None:
            self.mass = torch.as_tensor(np.asarray(masses, dtype=np.float64), device=self.device).contiguous()
            per = self.mass.ndim == 2
        self.desc = api.env_desc(n, leader_index, d0, t0, d_safe, quadratic_cost, False, per)

    def step(self, x, u, leader, gear=None):
        """x (B,2n), u (B,n), leader (B,2) [, gear (B,n) int32] CUDA tensors ->
        (x_new, cost, viol, err) CUDA tensors."""
        torch = self.torch
        B = x.shape[0]
        x_out = torch.empty_like(x)
        cost = torch.empty(B, dtype=torch.float64, device=self.device)
        viol = torch.empty(B, dtype=torch.uint8, device=self.device)
        err = torch.empty(B, dtype=torch.int32, device=self.device)
        api.rollout_step_device(self.desc, B, x, u, gear, self.mass, leader, x_out, cost, viol, err, ctx=self.ctx,
                                stream=torch.cuda.current_stream().cuda_stream)
        return x_out, cost, viol, err
