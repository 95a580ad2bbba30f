"""Import stub (test infrastructure only): forward-Euler discretisation used by
/root/reference/models.py:378 (dmpcrl==1.0.1 is not installed)."""
import numpy as np


def forward_euler(A, B, ts, c=None):
    Ad = np.eye(A.shape[0]) + ts * A
    Bd = ts * B
    if c is None:
        return Ad, Bd
    return Ad, Bd, ts * c
