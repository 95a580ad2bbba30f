"""Drop-in namespace for the reference's fleet_seq_mld.py (config 2, sequential)."""
from .agents import TrackingSequentialMldCoordinator, simulate as _simulate  # noqa: F401
from .mpc import LocalMpcGear, LocalMpcMld  # noqa: F401


def simulate(sim, save: bool = False, plot: bool = False, seed: int = 1, thread_limit=None, leader_index: int = 0,
             order_forwards: bool = True, **kw):
    """fleet_seq_mld.simulate (:443-548; same defaults: seed 1).  `order_forwards` is stored by the reference's
    coordinator (:330) and never read: the solve order is leader, vehicles in front, vehicles behind either way."""
    return _simulate(sim, "seq", seed=seed, leader_index=leader_index, save=save, **kw)
