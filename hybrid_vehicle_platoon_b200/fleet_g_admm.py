"""Drop-in for the reference's fleet_g_admm.py (config 3, "switching" ADMM): every agent holds a FIXED PWA
region sequence, so its local problem is a convex QP (fleet_g_admm.py:55-158); the sequences are
re-identified from a PWA roll-out of the current inputs between ADMM rounds.  The n local QPs of a round
are solved in one launch per distinct formulation by the compiled-MPC kernel (fixed_modes path).

The round logic lives in un-vendored third-party code in the reference (dmpcpwa GAdmmCoordinator.
g_admm_control, dmpcrl MpcAdmm); it is restated here from the reference's call sites
(fleet_g_admm.py:253-301) and the published g-ADMM scheme -- UNVERIFIED-3P."""
from __future__ import annotations

import numpy as np

from ._sim import collect, make_env_and_systems
from .agents import MldAgent
from .misc import Params, Sim
from .mpc import GAdmmLocalMpc as LocalMpc, solve_compiled_batch


def g_map(Adj: np.ndarray):
    """dmpcrl.core.admm.g_map: G[i] = sorted indices of agent i and its neighbours."""
    n = Adj.shape[0]
    return [sorted([i] + [j for j in range(n) if Adj[i, j] == 1]) for i in range(n)]


def _region_of(system, x, u):
    for r in range(len(system["S"])):
        if all(system["S"][r] @ x + system["R"][r] @ u <= system["T"][r] + 1e-9):
            return r
    raise RuntimeError("state outside every PWA region")


class GAdmmCoordinator(MldAgent):
    def __init__(self, local_mpcs, local_fixed_parameters, systems, G, Adj, rho, debug_plot=False,
                 admm_iters: int = 50) -> None:
        super().__init__(local_mpcs[0])
        self.n = len(local_mpcs)
        self.agents = [MldAgent(m) for m in local_mpcs]
        self.systems, self.G, self.Adj, self.rho, self.admm_iters = systems, G, Adj, rho, admm_iters
        self.nx_l, self.nu_l = 2, 1
        self.N = local_mpcs[0].N
        self.prev_sol = None
        self.prev_traj = None
        self.prev_sol_time = None

    def dynamics_rollout(self, x, u):
        """PWA roll-out of every vehicle under inputs u: trajectories (2, N+1) and region sequences."""
        trajs, seqs = [], []
        for i in range(self.n):
            sy = self.systems[i]
            xt = np.zeros((2, self.N + 1))
            xt[:, [0]] = x[i]
            seq = []
            for k in range(self.N):
                r = _region_of(sy, xt[:, [k]], u[i][:, [k]])
                seq.append(r)
                xt[:, [k + 1]] = sy["A"][r] @ xt[:, [k]] + sy["B"][r] @ u[i][:, [k]] + sy["c"][r]
            trajs.append(xt)
            seqs.append(seq)
        return trajs, seqs

    def g_admm_control(self, state, warm_start=None):
        """-> (u_opt (n,1), sol_list, error_flag, infeas_flag)."""
        n, N, G = self.n, self.N, self.G
        state = np.asarray(state, dtype=np.float64)
        x = [state[2 * i:2 * (i + 1), :] for i in range(n)]
        u = [np.array(w, dtype=np.float64).reshape(1, N) for w in warm_start] if warm_start is not None \
            else [np.zeros((1, N)) for _ in range(n)]
        y = [np.zeros((2 * len(G[i]), N + 1)) for i in range(n)]
        trajs, seqs = self.dynamics_rollout(x, u)
        z = [np.vstack([trajs[j] for j in G[i]]) for i in range(n)]
        run_time = 0.0
        mpcs = [a.mpc for a in self.agents]
        sol_list = [None] * n
        for it in range(self.admm_iters):
            for i in range(n):
                mpcs[i].set_sequence(seqs[i])
                mpcs[i].set_consensus(y[i], z[i])
            res = solve_compiled_batch(mpcs, x, raises=False)
            run_time += max(info["run_time"] for _, info in res)
            if any(info["status"] != 2 for _, info in res):
                return np.vstack([ui[:, [0]] for ui in u]), sol_list, True, True
            for i, (_, info) in enumerate(res):
                u[i] = info["u"]
                sol_list[i] = info
            aug = [mpcs[i].augmented() for i in range(n)]
            # z-update: average of every copy of vehicle j's trajectory; y-update with the new z
            zbar = []
            for j in range(n):
                copies = [aug[i][2 * G[i].index(j):2 * G[i].index(j) + 2, :] for i in range(n) if j in G[i]]
                zbar.append(sum(copies) / len(copies))
            for i in range(n):
                z[i] = np.vstack([zbar[j] for j in G[i]])
                y[i] = y[i] + self.rho * (aug[i] - z[i])
            trajs, seqs = self.dynamics_rollout(x, u)     # sequences of the next round follow the current inputs
        self.prev_sol, self.prev_traj, self.prev_sol_time = [ui.copy() for ui in u], trajs, run_time
        infeas = any((t[1, 1:] > 45.84 + 1e-6).any() or (t[1, 1:] < 3.94 - 1e-6).any() for t in trajs)
        return np.vstack([ui[:, [0]] for ui in u]), sol_list, False, infeas


class TrackingGAdmmCoordinator(GAdmmCoordinator):
    """fleet_g_admm.py:208-301."""

    def __init__(self, N: int, ep_len: int, leader_x: np.ndarray, local_mpcs, local_fixed_parameters, systems,
                 vehicles, G, Adj, rho: float, debug_plot: bool = False, admm_iters: int = 50) -> None:
        super().__init__(local_mpcs, local_fixed_parameters, systems, G, Adj, rho, debug_plot, admm_iters)
        self.N, self.ep_len, self.leader_x, self.vehicles = N, ep_len, leader_x, vehicles
        self.best_warm_starts: list = []
        self.solve_times: list = []
        self.node_counts = 0                 # the reference pickles 0 in the node-count slot (:435-437)

    def get_control(self, state):
        u, _, _, _ = self.g_admm_control(state)
        return u, {}

    def on_timestep_end(self, env, episode: int, timestep: int) -> None:
        self.set_leader_traj(self.leader_x[:, timestep:(timestep + self.N + 1)])

    def on_episode_start(self, env, episode: int, state) -> None:
        self.set_leader_traj(self.leader_x[:, 0:self.N + 1])

    def set_leader_traj(self, leader_traj):
        self.agents[0].mpc.set_leader_traj(leader_traj)      # first agent is the leader (:250-251)

    def g_admm_control(self, state, warm_start=None):
        state = np.asarray(state, dtype=np.float64)
        warm_start = [[self.vehicles[i].get_u_for_constant_vel(state[2 * i + 1, 0]) * np.ones((1, self.N))
                       for i in range(self.n)]]
        if self.prev_sol is not None:         # shifted previous solution (:266-272)
            warm_start.append([np.hstack((self.prev_sol[i][:, 1:], self.prev_sol[i][:, [-1]])) for i in range(self.n)])
        best_cost, best_control = float("inf"), np.zeros((self.n, 1))
        counter = 0
        self.best_warm_starts.append(counter)
        temp_solve_times = []
        prev_best = None
        for u in warm_start:
            counter += 1
            u_opt, sol_list, error_flag, infeas_flag = super().g_admm_control(state, warm_start=u)
            if not error_flag and not infeas_flag:
                cost = sum(sol_list[i]["cost"] for i in range(self.n))
                temp_solve_times.append(self.prev_sol_time)
            else:
                cost = float("inf")
            if cost < best_cost:
                best_cost, best_control = cost, u_opt
                self.best_warm_starts[-1] = counter
                prev_best = (self.prev_sol, self.prev_traj)
        if best_cost == float("inf"):
            self.solve_times.append(0.0)
            raise RuntimeError("No solution found for any of the warm starts")
        self.prev_sol, self.prev_traj = prev_best
        self.solve_times.append(max(temp_solve_times))
        return best_control, None, None, None


def simulate(sim: Sim, save: bool = False, plot: bool = False, seed: int = 1, admm_iters: int = 100, ep_len=None,
             env_class=None):
    """fleet_g_admm.simulate (:304-439)."""
    n, N = sim.n, sim.N
    if sim.vehicle_model_type != "pwa_gear":
        raise NotImplementedError()            # as the reference (:363-366)
    leader_x = sim.leader_trajectory.get_leader_trajectory()
    env, platoon, systems, ep_len = make_env_and_systems(sim, 0, ep_len, env_class, forward_real_ref=False)
    Adj = np.zeros((n, n))
    for i in range(n):
        if i > 0:
            Adj[i, i - 1] = 1
        if i < n - 1:
            Adj[i, i + 1] = 1
    G = g_map(Adj)
    mpcs = [LocalMpc(N=N, pwa_system=systems[i], spacing_policy=sim.spacing_policy, num_neighbours=len(G[i]) - 1,
                     my_index=G[i].index(i), leader=(i == 0)) for i in range(n)]
    agent = TrackingGAdmmCoordinator(N=N, ep_len=ep_len, leader_x=leader_x, local_mpcs=mpcs,
                                     local_fixed_parameters=[{} for _ in range(n)], systems=systems,
                                     vehicles=platoon.get_vehicles(), G=G, Adj=Adj, rho=LocalMpc.rho,
                                     admm_iters=admm_iters)
    agent.evaluate(env=env, episodes=1, seed=seed)
    return collect(env, agent, leader_x, f"switching_admm_{sim.id}_seed_{seed}.pkl", save, node_counts=0)
