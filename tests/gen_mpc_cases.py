"""Re-export: the seeded synthetic compiled-MPC workloads live in the package (bench.py uses them too)."""
from hybrid_vehicle_platoon_b200.synth_mpc import *  # noqa: F401,F403
from hybrid_vehicle_platoon_b200.synth_mpc import (ADMM, CENT, EDGES, EVENT, FRONT, GADMM, LEADER, LOCAL, TRAILER,  # noqa: F401
                                                   admm_cases, cent_cases, const_vel, event_cases, gadmm_cases,
                                                   platoon_states)
