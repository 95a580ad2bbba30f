"""Drop-in for the reference's fleet_cent_mld.py: centralized MLD MPC of the whole platoon (config 1).
MpcMldCent / MpcGearCent solve on the GPU (compiled-MPC kernel); TrackingCentralizedAgent and simulate()
follow fleet_cent_mld.py:80-213."""
from __future__ import annotations

import numpy as np

from ._sim import collect, make_env_and_systems
from .agents import MldAgent
from .misc import Params, Sim
from .mpc import MpcGearCent, MpcMldCent  # noqa: F401


class TrackingCentralizedAgent(MldAgent):
    def __init__(self, mpc: MpcMldCent, ep_len: int, N: int, leader_x: np.ndarray) -> None:
        self.solve_times = np.zeros((ep_len, 1))
        self.node_counts = np.zeros((ep_len, 1))
        self.bin_var_counts = np.zeros((ep_len, 1))
        self.N, self.leader_x = N, leader_x
        super().__init__(mpc)

    def on_timestep_end(self, env, episode: int, timestep: int) -> None:
        # time step starts from 1, so this sets the cost accurately for the next time-step
        self.mpc.set_leader_traj(self.leader_x[:, timestep:(timestep + self.N + 1)])
        self.solve_times[env.step_counter - 1, :] = self.run_time
        self.node_counts[env.step_counter - 1, :] = self.node_count
        self.bin_var_counts[env.step_counter - 1, :] = self.num_bin_vars
        return super().on_timestep_end(env, episode, timestep)

    def on_episode_start(self, env, episode: int, state) -> None:
        self.mpc.set_leader_traj(self.leader_x[:, 0:self.N + 1])
        return super().on_episode_start(env, episode, state)


def simulate(sim: Sim, save: bool = False, plot: bool = False, seed: int = 1, thread_limit=None,
             leader_index: int = 0, ep_len=None, env_class=None):
    """fleet_cent_mld.simulate (:104-213); returns the 7 result objects as a dict."""
    n, N = sim.n, sim.N
    leader_x = sim.leader_trajectory.get_leader_trajectory()
    env, platoon, systems, ep_len = make_env_and_systems(sim, leader_index, ep_len, env_class,
                                                         forward_quadratic=True)
    if sim.vehicle_model_type not in ("pwa_gear", "pwa_friction"):
        raise NotImplementedError(f"{sim.vehicle_model_type}: the non-convex nonlinear MPC is not built (SURVEY 8f)")
    mpc = MpcMldCent(n, N, systems, spacing_policy=sim.spacing_policy, leader_index=leader_index,
                     quadratic_cost=sim.quadratic_cost, thread_limit=thread_limit,
                     real_vehicle_as_reference=sim.real_vehicle_as_reference)
    agent = TrackingCentralizedAgent(mpc, ep_len, N, leader_x)
    agent.evaluate(env=env, episodes=1, seed=seed, open_loop=sim.open_loop)
    return collect(env, agent, leader_x, f"cent_{sim.id}_seed_{seed}.pkl", save)
