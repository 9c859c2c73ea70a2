"""Update-cost functions (API compatibility with ``python/repair/costs.py:25-78``).

They are only consulted by the pmf-weighting / nearest-value modes, which are outside this
version's hot path; the classes exist so that ``setUpdateCostFunction`` type-checks like the
reference and the Levenshtein distance is available to callers.
"""
from abc import ABCMeta, abstractmethod
from typing import List, Optional, Union


class UpdateCostFunction(metaclass=ABCMeta):

    def __init__(self, targets: List[str] = []) -> None:
        self.targets: List[str] = targets

    @abstractmethod
    def _compute_impl(self, x: Union[str, int, float], y: Union[str, int, float]) -> Optional[float]:
        pass

    def compute(self, x: Union[str, int, float], y: Union[str, int, float]) -> Optional[float]:
        return self._compute_impl(x, y)


class Levenshtein(UpdateCostFunction):

    def __init__(self, targets: List[str] = []) -> None:
        UpdateCostFunction.__init__(self, targets)

    def __str__(self) -> str:
        return "{}(targets={})".format(self.__class__.__name__, ",".join(self.targets))

    def _compute_impl(self, x, y):
        a, b = str(x), str(y)
        if a == b:
            return 0.0
        prev = list(range(len(b) + 1))
        for i, ca in enumerate(a, 1):
            cur = [i]
            for j, cb in enumerate(b, 1):
                cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
            prev = cur
        return float(prev[-1])


class UserDefinedUpdateCostFunction(UpdateCostFunction):

    def __init__(self, f, targets: List[str] = []) -> None:
        UpdateCostFunction.__init__(self, targets)
        if not callable(f):
            raise ValueError("`f` should be callable")
        ret = f("x", "y")
        if type(ret) is not float:
            raise ValueError("`f` should return a float cost value")
        self.f = f

    def __str__(self) -> str:
        return "{}(targets={})".format(self.__class__.__name__, ",".join(self.targets))

    def _compute_impl(self, x, y):
        return self.f(str(x), str(y))
