"""hybrid_vehicle_platoon_b200 -- B200-native (sm_100a) batched hybrid-MPC hot path for
Kevindqz/hybrid-vehicle-platoon: rollout/stage-cost kernel and per-vehicle MIQP
branch-and-bound kernel behind a C ABI (include/hvp.h), with reference-shaped Python classes.
"""
from . import _lib  # noqa: F401
from ._lib import FRONT, LEADER, TRAILER, Context, default_context  # noqa: F401
from .api import (  # noqa: F401
    CompiledMpc, env_desc, local_desc, local_miqp, local_miqp_device, microbench_fp64, rollout_step,
    rollout_step_device,
)

__version__ = "0.1.0"
from .agents import (  # noqa: F401,E402
    MldAgent, MonitorEpisodes, TimeLimit, TrackingDecentMldCoordinator, TrackingSequentialMldCoordinator,
    simulate,
)
from .env import BatchedPlatoonEnv, PlatoonEnv  # noqa: F401,E402
from .misc import (  # noqa: F401,E402
    ConstantSpacingPolicy, ConstantTimePolicy, ConstantVelocityLeaderTrajectory, Params, Sim, Sim_n_task_1,
    Sim_n_task_2, StopAndGoLeaderTrajectory,
)
from .models import Platoon, PwaFrictionVehicle, PwaGearVehicle, Vehicle  # noqa: F401,E402
from .mpc import (  # noqa: F401,E402
    EventLocalMpc, GAdmmLocalMpc, LocalMpcADMM, LocalMpcGear, LocalMpcMld, MpcGearCent, MpcMldCent,
    eval_compiled_batch, set_solver_options, solve_compiled_batch, solve_local_batch,
)
from . import (  # noqa: F401,E402
    fleet_cent_mld, fleet_decent_mld, fleet_event_based, fleet_g_admm, fleet_naive_admm, fleet_seq_mld,
)
