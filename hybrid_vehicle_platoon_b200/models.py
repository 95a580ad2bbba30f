"""Host-side vehicle / platoon model objects with the reference's interface (models.py:54-556):
constants, PWA-gear system dictionaries {S,R,T,A,B,c,D,E,F,G}, gear lookup and platoon stepping.
Platoon.step_platoon runs on the GPU (rollout kernel); there is no CPU stepping code here."""
from __future__ import annotations

import warnings

import numpy as np


class Vehicle:
    nx_l = 2
    nu_l = 1
    c_fric = 0.5
    mu = 0.01
    grav = 9.8
    b = [4057, 2945, 2116, 1607, 1166, 838]
    vl = [3.94, 5.43, 7.56, 9.96, 13.70, 19.10]
    vh = [9.46, 13.04, 18.15, 23.90, 32.93, 45.84]
    v_min, v_max = vl[0], vh[-1]
    u_min, u_max = -1.0, 1.0
    p_min, p_max = 0.0, 10000.0

    def __init__(self, m: float = 800) -> None:
        self.m = m


class PwaGearVehicle(Vehicle):
    """PWA approximation of gears and friction: 7 velocity regions (models.py:390-556)."""

    beta = (3 * Vehicle.c_fric * Vehicle.v_max**2) / 16
    alpha = Vehicle.v_max / 2
    c1 = beta / alpha
    c2 = (Vehicle.c_fric * Vehicle.v_max**2 - beta) / (Vehicle.v_max - alpha)
    d = beta - alpha * ((Vehicle.c_fric * Vehicle.v_max**2 - beta) / (Vehicle.v_max - alpha))

    def __init__(self, m: float = 800) -> None:
        super().__init__(m)
        self.v_gear_lim = [(self.vh[i] - self.vl[i]) / 2 + self.vl[i] for i in range(1, 6)]
        self.system = self._build(m)

    def _build(self, mass):
        lim = self.v_gear_lim
        edges = [lim[0], lim[1], lim[2], self.alpha, lim[3], lim[4]]
        S = [np.array([[0, 1], [0, 0]])] + [np.array([[0, 1], [0, -1]]) for _ in range(5)] + [np.array([[0, 0], [0, -1]])]
        T = [np.array([[edges[0]], [0]])]
        T += [np.array([[edges[i]], [-edges[i - 1]]]) for i in range(1, 6)]
        T += [np.array([[0], [-edges[5]]])]
        R = [np.zeros((2, 1)) for _ in range(7)]
        gear_of_region = [0, 1, 2, 3, 3, 4, 5]
        A = [np.array([[0, 1], [0, -(self.c1 if r < 4 else self.c2) / mass]]) for r in range(7)]
        B = [np.array([[0], [self.b[gear_of_region[r]] / mass]]) for r in range(7)]
        c = [np.array([[0], [-self.mu * self.grav - (0 if r < 4 else self.d / mass)]]) for r in range(7)]
        D = np.array([[1, 0], [-1, 0], [0, 1], [0, -1]])
        E = np.array([[self.p_max], [-self.p_min], [self.v_max], [-self.v_min]])
        F = np.array([[1], [-1]])
        G = np.array([[self.u_max], [-self.u_min]])
        return {"S": S, "R": R, "T": T, "A": A, "B": B, "c": c, "D": D, "E": E, "F": F, "G": G}

    def get_discrete_system(self, ts: float):
        """Forward-Euler discretisation (models.py:370-387): A_d = I + ts A, B_d = ts B, c_d = ts c."""
        sysd = dict(self.system)
        sysd["A"] = [np.eye(2) + ts * A for A in self.system["A"]]
        sysd["B"] = [ts * B for B in self.system["B"]]
        sysd["c"] = [ts * c for c in self.system["c"]]
        return sysd

    def get_gear_from_velocity(self, v: float) -> int:
        lim = self.v_gear_lim
        for i in range(4):
            if lim[i] <= v < lim[i + 1]:
                return i + 2
        if v < lim[0]:
            if v < self.v_min:
                warnings.warn(f"Velocity {v} is below min {self.v_min}, using first gear but result will be approximate.")
            return 1
        if v > self.v_max:
            warnings.warn(f"Velocity {v} is above max {self.v_max}, using last gear but result will be approximate.")
        return 6

    def get_u_for_constant_vel(self, v: float) -> float:
        """Throttle keeping v constant under the PWA dynamics (models.py:537-556)."""
        for j in range(7):
            S, T = self.system["S"][j], self.system["T"][j]
            if all(S @ np.array([[0], [v]]) <= T + np.array([[0], [1e-4]])):
                return (1 / self.system["B"][j][1, 0]) * (-self.system["A"][j][1, 1] * v - self.system["c"][j][1, 0])
        raise RuntimeError("Didn't find any PWA region for the given speed!")


class PwaFrictionVehicle(Vehicle):
    """PWA approximation of the friction only: 2 velocity regions, input = traction force with
    B = 1/m; used with the discrete gears of MpcGear (models.py:272-387)."""

    beta, alpha, c1, c2, d = PwaGearVehicle.beta, PwaGearVehicle.alpha, PwaGearVehicle.c1, PwaGearVehicle.c2, PwaGearVehicle.d

    def __init__(self, m: float = 800) -> None:
        super().__init__(m)
        self.system = self._build(m)

    def _build(self, mass):
        S = [np.array([[0, 1], [0, 0]]), np.array([[0, 0], [0, -1]])]
        T = [np.array([[self.alpha], [0]]), np.array([[0], [-self.alpha]])]
        R = [np.zeros((2, 1)), np.zeros((2, 1))]
        A = [np.array([[0, 1], [0, -self.c1 / mass]]), np.array([[0, 1], [0, -self.c2 / mass]])]
        B = [np.array([[0], [1 / mass]]), np.array([[0], [1 / mass]])]
        c = [np.array([[0], [-self.mu * self.grav]]), np.array([[0], [-self.mu * self.grav - self.d / mass]])]
        D = np.array([[1, 0], [-1, 0], [0, 1], [0, -1]])
        E = np.array([[self.p_max], [-self.p_min], [self.v_max], [-self.v_min]])
        F = np.array([[1], [-1]])
        G = np.array([[self.u_max], [-self.u_min]])
        return {"S": S, "R": R, "T": T, "A": A, "B": B, "c": c, "D": D, "E": E, "F": F, "G": G}

    get_discrete_system = PwaGearVehicle.get_discrete_system

    def get_gear_from_velocity(self, v: float) -> int:
        """Vehicle.get_gear_from_velocity (models.py:161-174): first gear whose OPEN window contains v."""
        if v < self.v_min or v > self.v_max:
            warnings.warn(f"Velocity {v} is not within bounds {self.v_min}/{self.v_max}")
            if v < self.v_min:
                return 1
            if v > self.v_max:
                return 6
        for i in range(len(self.b)):
            if self.vl[i] < v < self.vh[i]:
                return i + 1
        raise ValueError(f"No gear found for velocity {v}")

    def get_u_for_constant_vel(self, v: float, j: int) -> float:
        """Throttle keeping v constant in gear j under the PWA-friction dynamics (models.py:334-368)."""
        if j < 1 or j > 6:
            raise ValueError(f"{j} is not a valid gear.")
        j -= 1
        if not ((v < self.v_min and j == 0) or (v > self.v_max and j == 5)) and (v < self.vl[j] or v > self.vh[j]):
            raise ValueError(f"Velocity {v} is not valid for gear {j + 1}")
        for i in range(2):
            S, T = self.system["S"][i], self.system["T"][i]
            if all(S @ np.array([[0], [v]]) <= T + np.array([[0], [1e-4]])):
                return (1 / (self.b[j] * self.system["B"][i][1, 0])) * (
                    -self.system["A"][i][1, 1] * v - self.system["c"][i][1, 0])
        raise RuntimeError("Didn't find any PWA region for the given speed!")


def model_of_pwa_system(system: dict):
    """(model id, mass) of a system dict (ours or the reference's): 0 = pwa_gear (7 regions),
    1 = pwa_friction (2 regions; solved with the six discrete gears of MpcGear)."""
    nreg = len(system["A"])
    if nreg == 7:
        return 0, mass_of_pwa_system(system)
    if nreg == 2:
        m = 1.0 / float(system["B"][0][1, 0])
        ref = PwaFrictionVehicle(m).get_discrete_system(1.0)
        if not all(np.allclose(system[k][r], ref[k][r], rtol=1e-9, atol=1e-12) for k in ("A", "B", "c", "T") for r in range(2)):
            raise NotImplementedError("only the pwa_friction model with ts = 1 is implemented on the GPU path")
        return 1, m
    raise NotImplementedError("GPU MPC needs a pwa_gear (7 regions) or pwa_friction (2 regions) system dict")


def mass_of_pwa_system(system: dict) -> float:
    """Recover the vehicle mass from a pwa_gear system dict (ours or the reference's) and check
    that the dict really is that model (the GPU path implements pwa_gear only)."""
    try:
        ok = len(system["A"]) == 7 and len(system["S"]) == 7
        m = Vehicle.b[0] / float(system["B"][0][1, 0])
    except Exception as e:  # pragma: no cover
        raise NotImplementedError("GPU MPC needs a pwa_gear system dict {S,R,T,A,B,c,D,E,F,G}") from e
    ref = PwaGearVehicle(m).get_discrete_system(1.0)
    if not ok or not all(np.allclose(system[k][r], ref[k][r], rtol=1e-9, atol=1e-12)
                         for k in ("A", "B", "c", "T") for r in range(7)):
        raise NotImplementedError(
            "only the pwa_gear model with ts = 1 is implemented on the GPU path "
            "(pwa_friction + discrete gears and the nonlinear model are SURVEY.md 8f rows)")
    return m


class Platoon:
    nx_l = Vehicle.nx_l
    nu_l = Vehicle.nu_l

    def __init__(self, n: int, vehicle_type: str = "pwa_gear", masses=None) -> None:
        cls = {"pwa_gear": PwaGearVehicle, "pwa_friction": PwaFrictionVehicle}.get(vehicle_type)
        if cls is None:
            raise NotImplementedError(f"vehicle_type {vehicle_type!r}: 'pwa_gear' and 'pwa_friction' are implemented "
                                      "(the non-convex 'nonlinear' MPC model is SURVEY.md 8f rank 4)")
        if masses is not None and len(masses) != n:
            raise ValueError(f"Required {n} vehicles masses. Got {len(masses)}.")
        self.n = n
        self.vehicles = [cls(m=masses[i]) if masses is not None else cls() for i in range(n)]

    @property
    def masses(self) -> np.ndarray:
        return np.array([float(v.m) for v in self.vehicles])

    def get_vehicles(self):
        return self.vehicles

    def get_gear_from_vehicle_velocity(self, i: int, v: float) -> int:
        return self.vehicles[i].get_gear_from_velocity(v)

    def get_vehicle_system_dicts(self, ts: float):
        return [v.get_discrete_system(ts) for v in self.vehicles]

    def step_platoon(self, x: np.ndarray, u: np.ndarray, j: np.ndarray, ts: float = 1.0) -> np.ndarray:
        """models.py:236-257 on the GPU (ts = 1 only); raises like the reference on range errors."""
        from .env import raise_rollout_error
        from .api import rollout_step
        if x.shape != (2 * self.n, 1) or u.shape != (self.n, 1) or j.shape != (self.n, 1):
            raise ValueError("Dimension error in x, u, or j.")
        if ts != 1:
            raise NotImplementedError("the rollout kernel implements the reference's ts = 1")
        xo, _, _, err = rollout_step(x.reshape(1, -1), u.reshape(1, -1), j.reshape(1, -1).astype(np.int32),
                                     self.masses, np.zeros((1, 2)))
        raise_rollout_error(int(err[0]))
        return xo.reshape(-1, 1)
