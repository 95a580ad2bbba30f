"""Import stub (test infrastructure only): dmpcpwa==0.0.2 is un-vendored third-party code
(requirements.txt:9).  The reference's coordinators subclass its MldAgent; here that name resolves to
the repo's behaviour-compatible restatement, so that /root/reference/fleet_*.py
import and run UNMODIFIED (tests/golden/make_fleet_golden.py)."""
from hybrid_vehicle_platoon_b200.agents import MldAgent  # noqa: F401
