"""Import stub (test infrastructure only): mpcrl.wrappers.envs.MonitorEpisodes -> the repo's restatement
(buffers pinned by fleet_cent_mld.py:185-192)."""
from hybrid_vehicle_platoon_b200.agents import MonitorEpisodes  # noqa: F401
