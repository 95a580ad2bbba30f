"""Agent loop, env wrappers and fleet coordinators on top of the GPU controllers.

The reference takes these from un-vendored packages (dmpcpwa.agents.mld_agent.MldAgent,
gymnasium.wrappers.TimeLimit, mpcrl.wrappers.envs.MonitorEpisodes); their behaviour is restated
here from the reference's call sites.  The coordinators mirror
fleet_decent_mld.py:285-455 and fleet_seq_mld.py:296-440; the decentralized one solves the n
local MIQPs of a timestep in ONE kernel launch (they are independent), the sequential one keeps the
reference's data-dependent chain (vehicle i needs i-1's fresh prediction)."""
from __future__ import annotations

from collections import deque

import numpy as np

from .misc import Params
from .mpc import LocalMpcGear, LocalMpcMld, _CompiledController, solve_compiled_batch, solve_local_batch


# ---- env wrappers ---------------------------------------------------------------------------
class _Wrapper:
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):          # forward attribute access to the base env
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        e = self.env
        while hasattr(e, "env"):
            e = e.env
        return e

    def reset(self, **kw):
        return self.env.reset(**kw)

    def step(self, action):
        return self.env.step(action)


class TimeLimit(_Wrapper):
    """Truncates the episode after max_episode_steps (gymnasium.wrappers.TimeLimit)."""

    def __init__(self, env, max_episode_steps: int):
        super().__init__(env)
        self._max, self._t = max_episode_steps, 0

    def reset(self, **kw):
        self._t = 0
        return self.env.reset(**kw)

    def step(self, action):
        obs, r, terminated, truncated, info = self.env.step(action)
        self._t += 1
        if self._t >= self._max:
            truncated = True
        return obs, r, terminated, truncated, info


class MonitorEpisodes(_Wrapper):
    """Records observations (T+1,2n,1), actions (T,m,1), rewards (T,1,1) per finished episode in
    deques `observations`, `actions`, `rewards` (mpcrl.wrappers.envs.MonitorEpisodes; read at
    fleet_cent_mld.py:185-192)."""

    def __init__(self, env):
        super().__init__(env)
        self.observations, self.actions, self.rewards = deque(), deque(), deque()
        self.ep_observations, self.ep_actions, self.ep_rewards = [], [], []

    def reset(self, **kw):
        obs, info = self.env.reset(**kw)
        self.ep_observations, self.ep_actions, self.ep_rewards = [np.array(obs, dtype=np.float64)], [], []
        return obs, info

    def step(self, action):
        obs, r, terminated, truncated, info = self.env.step(action)
        self.ep_observations.append(np.array(obs)); self.ep_actions.append(np.array(action))
        self.ep_rewards.append(np.array(r))
        if terminated or truncated:
            self.observations.append(np.array(self.ep_observations))
            self.actions.append(np.array(self.ep_actions))
            self.rewards.append(np.array(self.ep_rewards))
            self.ep_observations, self.ep_actions, self.ep_rewards = [], [], []
        return obs, r, terminated, truncated, info


# ---- agent ----------------------------------------------------------------------------------
class MldAgent:
    """dmpcpwa.agents.mld_agent.MldAgent as the reference uses it."""

    def __init__(self, mpc) -> None:
        self.mpc = mpc
        self.x_pred = self.u_pred = self.cost_pred = None
        self.run_time = self.node_count = self.num_bin_vars = None

    def evaluate(self, env, episodes: int, seed=None, open_loop: bool = False):
        returns = np.zeros(episodes)
        seeds = np.random.SeedSequence(seed).generate_state(episodes)   # model_validation.py:76
        for ep in range(episodes):
            state, _ = env.reset(seed=int(seeds[ep]))
            truncated = terminated = False
            timestep = 0
            self.on_episode_start(env, ep, state)
            if open_loop:
                _, info = self.get_control(state)
                actions = info["u"]
            while not (truncated or terminated):
                action = actions[:, [timestep]] if open_loop else self.get_control(state)[0]
                state, r, truncated, terminated, _ = env.step(action)
                returns[ep] += float(np.asarray(r).ravel()[0])
                timestep += 1
                self.on_timestep_end(env, ep, timestep)
            self.on_episode_end(env, ep, returns[ep])
        return returns

    def _store(self, info):
        self.x_pred, self.u_pred, self.cost_pred = info["x"], info["u"], info["cost"]
        self.run_time, self.node_count, self.num_bin_vars = info["run_time"], info["nodes"], info["bin_vars"]

    def get_control(self, state):
        u, info = self.mpc.solve_mpc(state)
        self._store(info)
        return u, info

    def get_predicted_state(self, shifted: bool = False):
        if self.x_pred is None:
            return None
        if shifted:                       # drop column 0, repeat the last one
            return np.concatenate((self.x_pred[:, 1:], self.x_pred[:, -1:]), axis=1)
        return self.x_pred.copy()

    def get_predicted_cost(self):
        return self.cost_pred

    def on_episode_start(self, env, episode, state): pass
    def on_timestep_end(self, env, episode, timestep): pass
    def on_episode_end(self, env, episode, rewards): pass


def _solve_batch(mpcs, states):
    """All local problems of one wave in as few launches as possible."""
    comp = [i for i, m in enumerate(mpcs) if isinstance(m, _CompiledController)]
    if len(comp) == len(mpcs):
        return solve_compiled_batch(mpcs, states)
    if not comp:
        return solve_local_batch(mpcs, states)
    out = [None] * len(mpcs)                      # mixed controller types: one batch per type
    rest = [i for i in range(len(mpcs)) if i not in set(comp)]
    for idx, fn in ((comp, solve_compiled_batch), (rest, solve_local_batch)):
        for i, r in zip(idx, fn([mpcs[i] for i in idx], [states[i] for i in idx])):
            out[i] = r
    return out


def _stack_controls(u, nu_l=1):
    """[u_g ; gear] per vehicle -> all throttles, then all gears (fleet_decent_mld.py:317-326)."""
    if u[0].shape[0] > nu_l:
        return np.vstack((np.vstack([ui[:nu_l, :] for ui in u]), np.vstack([ui[nu_l:, :] for ui in u])))
    return np.vstack(u)


def _extrapolate_constant_vel(p, v, N, ts):
    """fleet_decent_mld.py:421-428: p_{k+1} = p_k + ts*v_k (sequential sums), v constant."""
    x = np.zeros((2, N + 1))
    x[0, 0], x[1, 0] = p, v
    for k in range(N):
        x[0, k + 1] = x[0, k] + ts * x[1, k]
        x[1, k + 1] = x[1, k]
    return x


def _extrapolate_two_point(p, v, v_prev, N, ts, sat_steps=None):
    """fleet_decent_mld.py:430-455: velocity keeps changing by dv = v - v_prev (for the first
    floor(N/2) steps only in the saturated variant)."""
    x = np.zeros((2, N + 1))
    x[0, 0], x[1, 0] = p, v
    dv = v - v_prev
    for k in range(N):
        x[0, k + 1] = x[0, k] + ts * x[1, k]
        x[1, k + 1] = x[1, k] + (dv if (sat_steps is None or k < sat_steps) else 0.0)
    return x


class TrackingDecentMldCoordinator(MldAgent):
    """fleet_decent_mld.py:285-455 with the n independent local solves batched into one launch."""

    def __init__(self, local_mpcs, ep_len: int, N: int, leader_x: np.ndarray, ts: float, leader_index: int = 0,
                 velocity_estimator="none") -> None:
        super().__init__(local_mpcs[0])
        self.n, self.ep_len, self.ts, self.N = len(local_mpcs), ep_len, ts, N
        self.leader_x, self.leader_index, self.velocity_estimator = leader_x, leader_index, velocity_estimator
        self.nx_l, self.nu_l = 2, 1
        self.agents = [MldAgent(m) for m in local_mpcs]
        self.solve_times = np.zeros((ep_len, 1))
        self.node_counts = np.zeros((ep_len, 1))

    def get_control(self, state):
        x_l = np.split(np.asarray(state, dtype=np.float64), self.n, axis=0)
        res = _solve_batch([a.mpc for a in self.agents], x_l)
        for a, (_, info) in zip(self.agents, res):
            a._store(info)
        return _stack_controls([u for u, _ in res]), {}

    def on_timestep_end(self, env, episode, timestep) -> None:
        self.agents[self.leader_index].mpc.set_leader_x(self.leader_x[:, timestep:timestep + self.N + 1])
        self.observe_states(env, timestep)
        # all agents are assumed to have operated in parallel: max time, max node count
        self.solve_times[env.step_counter - 1, :] = max(a.run_time for a in self.agents)
        self.node_counts[env.step_counter - 1, :] = max(a.node_count for a in self.agents)

    def on_episode_start(self, env, episode, state) -> None:
        self.agents[self.leader_index].mpc.set_leader_x(self.leader_x[:, 0:self.N + 1])
        self.observe_states(env, timestep=0)

    def _predict(self, env, j):
        p, v = float(env.x[2 * j, 0]), float(env.x[2 * j + 1, 0])
        if self.velocity_estimator == "two_point":
            return _extrapolate_two_point(p, v, float(env.get_previous_state()[2 * j + 1, 0]), self.N, self.ts)
        if self.velocity_estimator == "sat":
            return _extrapolate_two_point(p, v, float(env.get_previous_state()[2 * j + 1, 0]), self.N, self.ts,
                                          sat_steps=self.N // 2)
        return _extrapolate_constant_vel(p, v, self.N, self.ts)      # incl. Q9: any other value

    def observe_states(self, env, timestep):
        for i in range(self.n):
            if i != 0:
                self.agents[i].mpc.set_x_front(self._predict(env, i - 1))
            if i != self.n - 1:
                self.agents[i].mpc.set_x_back(self._predict(env, i + 1))


class TrackingSequentialMldCoordinator(MldAgent):
    """fleet_seq_mld.py:296-440: leader first, then outward, each vehicle using the fresh prediction
    of the neighbour already solved and the shifted previous prediction of the other."""

    def __init__(self, local_mpcs, ep_len: int, N: int, leader_x: np.ndarray, ts: float, leader_index: int = 0,
                 order_forwards: bool = True) -> None:
        super().__init__(local_mpcs[0])
        self.n = len(local_mpcs)
        self.agents = [MldAgent(m) for m in local_mpcs]
        self.solve_times = np.zeros((ep_len, 1))
        self.node_counts = np.zeros((ep_len, 1))
        self.leader_x, self.ep_len, self.ts, self.N = leader_x, ep_len, ts, N
        self.leader_index, self.forwards = leader_index, order_forwards
        self.nx_l, self.nu_l = 2, 1

    def _shifted(self, j):
        x = self.agents[j].get_predicted_state(shifted=True)
        if x is not None:
            x[0, -1] = x[0, -2] + self.ts * x[1, -1]
        return x

    def get_control(self, state):
        x_l = np.split(np.asarray(state, dtype=np.float64), self.n, axis=0)
        u = [None] * self.n
        li, n, ag = self.leader_index, self.n, self.agents
        if li != 0:
            xa = self._shifted(li - 1)
            if xa is not None:
                ag[li].mpc.set_x_front(xa)
        if li != n - 1:
            xb = self._shifted(li + 1)
            if xb is not None:
                ag[li].mpc.set_x_back(xb)
        u[li], _ = ag[li].get_control(x_l[li])
        for i in range(li - 1, -1, -1):                 # vehicles in front of the leader
            if i != 0:
                xa = self._shifted(i - 1)
                if xa is not None:
                    ag[i].mpc.set_x_front(xa)
            ag[i].mpc.set_x_back(ag[i + 1].get_predicted_state(shifted=False))
            u[i], _ = ag[i].get_control(x_l[i])
        for i in range(li + 1, n):                      # vehicles behind the leader
            ag[i].mpc.set_x_front(ag[i - 1].get_predicted_state(shifted=False))
            if i != n - 1:
                xb = self._shifted(i + 1)
                if xb is not None:
                    ag[i].mpc.set_x_back(xb)
            u[i], _ = ag[i].get_control(x_l[i])
        return _stack_controls(u), {}

    def on_timestep_end(self, env, episode, timestep) -> None:
        self.agents[self.leader_index].mpc.set_leader_x(self.leader_x[:, timestep:timestep + self.N + 1])
        self.solve_times[env.step_counter - 1, :] = sum(a.run_time for a in self.agents)   # solved in series
        self.node_counts[env.step_counter - 1, :] = max(a.node_count for a in self.agents)

    def on_episode_start(self, env, episode, state) -> None:
        self.agents[self.leader_index].mpc.set_leader_x(self.leader_x[:, 0:self.N + 1])
        for i in range(self.n):
            if i != 0:
                self.agents[i].mpc.set_x_front(
                    _extrapolate_constant_vel(float(env.x[2 * (i - 1), 0]), float(env.x[2 * (i - 1) + 1, 0]), self.N, self.ts))
            if i != self.n - 1:
                self.agents[i].mpc.set_x_back(
                    _extrapolate_constant_vel(float(env.x[2 * (i + 1), 0]), float(env.x[2 * (i + 1) + 1, 0]), self.N, self.ts))


# ---- simulate() of the fleet scripts -----------------------------------------------------------
def simulate(sim, controller: str = "decent", seed: int = 2, leader_index: int = 0, velocity_estimator="none",
             mpc_class=LocalMpcMld, env_class=None, save: bool = False, ep_len=None):
    """fleet_decent_mld.simulate (:458-559) / fleet_seq_mld.simulate (:443-548) on the GPU classes.
    Returns the reference's 7 result objects as a dict (X, U, R, solve_times, node_counts,
    violations, leader_x); `save=True` pickles them under the reference's file-name scheme."""
    from .env import PlatoonEnv
    from .models import Platoon
    env_class = env_class or PlatoonEnv
    if sim.vehicle_model_type == "pwa_friction" and mpc_class is LocalMpcMld:
        mpc_class = LocalMpcGear              # fleet_decent_mld.py:497-505: the model type picks the MPC class
    n, N, ts = sim.n, sim.N, Params.ts
    ep_len = ep_len or sim.ep_len
    leader_x = sim.leader_trajectory.get_leader_trajectory()
    platoon = Platoon(n, vehicle_type=sim.vehicle_model_type, masses=sim.masses)
    systems = platoon.get_vehicle_system_dicts(ts)
    env = MonitorEpisodes(TimeLimit(env_class(
        n=n, platoon=platoon, leader_trajectory=sim.leader_trajectory, spacing_policy=sim.spacing_policy,
        start_from_platoon=sim.start_from_platoon, real_vehicle_as_reference=sim.real_vehicle_as_reference,
        ep_len=ep_len, leader_index=leader_index), max_episode_steps=ep_len))
    mpcs = [mpc_class(N, systems[i], sim.spacing_policy, is_front=(i == 0), is_leader=(i == leader_index),
                      is_trailer=(i == n - 1), real_vehicle_as_reference=sim.real_vehicle_as_reference)
            for i in range(n)]
    if controller == "decent":
        agent = TrackingDecentMldCoordinator(mpcs, ep_len=ep_len, N=N, leader_x=leader_x, ts=ts,
                                             velocity_estimator=velocity_estimator, leader_index=leader_index)
        fname = f"decent_vest_{velocity_estimator}_{sim.id}_seed_{seed}.pkl"
    elif controller == "seq":
        agent = TrackingSequentialMldCoordinator(mpcs, ep_len=ep_len, N=N, leader_x=leader_x, ts=ts,
                                                 leader_index=leader_index)
        fname = f"seq_{sim.id}_seed_{seed}.pkl"
    else:
        raise ValueError(f"unknown controller {controller!r}")
    agent.evaluate(env=env, episodes=1, seed=seed)
    X, U, R = env.observations[0].squeeze(), env.actions[0].squeeze(), env.rewards[0]
    out = dict(X=X, U=U, R=R, solve_times=agent.solve_times, node_counts=agent.node_counts,
               violations=env.unwrapped.viol_counter[-1], leader_x=leader_x)
    if save:
        import pickle
        with open(fname, "wb") as f:
            for k in ("X", "U", "R", "solve_times", "node_counts", "violations", "leader_x"):
                pickle.dump(out[k], f)
    return out
