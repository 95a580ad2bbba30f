"""Import stub (test infrastructure only): gymnasium.wrappers.TimeLimit -> the repo's restatement."""
from hybrid_vehicle_platoon_b200.agents import TimeLimit  # noqa: F401
