"""Re-export: the seeded synthetic local-MIQP workload lives in the package (bench.py uses it too)."""
from hybrid_vehicle_platoon_b200.synth_local import *  # noqa: F401,F403
from hybrid_vehicle_platoon_b200.synth_local import EDGES, FRONT, LEADER, TRAILER, platoon_local_problems  # noqa: F401
