"""Import stub (test infrastructure only), see mpc_mld.py."""
from .mpc_mld import MpcMld


class MpcMldCentDecup(MpcMld):
    pass
