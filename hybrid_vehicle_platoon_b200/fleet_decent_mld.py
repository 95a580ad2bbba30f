"""Drop-in namespace for the reference's fleet_decent_mld.py (config 2, decentralized)."""
from .agents import TrackingDecentMldCoordinator, simulate as _simulate  # noqa: F401
from .mpc import LocalMpcGear, LocalMpcMld  # noqa: F401


def simulate(sim, save: bool = False, plot: bool = False, seed: int = 2, thread_limit=None, leader_index: int = 0,
             velocity_estimator="none", **kw):
    """fleet_decent_mld.simulate (:458-559)."""
    return _simulate(sim, "decent", seed=seed, leader_index=leader_index, velocity_estimator=velocity_estimator,
                     save=save, **kw)
