"""Import stub (test infrastructure only): lets /root/reference/env.py import
without the real gymnasium package. Not product code."""
from typing import Any, Generic, TypeVar

O = TypeVar("O")
A = TypeVar("A")


class Env(Generic[O, A]):
    def reset(self, *, seed: Any = None, options: Any = None) -> None:
        return None
