"""Import stub (test infrastructure only): empty gurobipy so that
/root/reference/models.py imports; only Vehicle.dyn (models.py:135) would use it."""


class GRB:
    MINIMIZE = 1
    BINARY = "B"
