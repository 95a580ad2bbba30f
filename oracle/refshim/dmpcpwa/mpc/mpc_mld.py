"""Import stub (test infrastructure only): base class name for the reference's `class LocalMpcMld(MpcMld)`
definitions.  Those classes are never instantiated in the golden runs -- the module-level names are
re-bound to the repo's controllers before simulate() is called."""


class MpcMld:
    def __init__(self, *a, **k):
        raise RuntimeError("dmpcpwa.MpcMld is a stub: gurobipy/dmpcpwa are not installable here")
