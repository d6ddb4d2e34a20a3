"""Host-side mirror of evolution/tree/SCMigrationModel.java and TypeSet.java.

Only Input resolution lives here (which of the three rate-matrix encodings, scale factors,
BSSVS flags, forward/backward conversion -- SCMigrationModel.java:119-133, 252-303, 324-345,
373-375); that is negligible O(d^2) host work done once per parameter change, and the result
(dense m[d x d], N[d]) is what crosses the C-ABI.  Q, R, mu, R^k and expm are computed by the
CUDA library (csrc/k3_unif.cu).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np



class TypeSet:
    """evolution/tree/TypeSet.java: ordered (sorted) set of type names; nTypes >= 2."""

    def __init__(self, *names: str):
        self._names = sorted(set(names))

    def getNTypes(self) -> int:
        return max(len(self._names), 2)            # TypeSet.java:59-61

    def getTypeIndex(self, name: str) -> int:
        return self._names.index(name)

    def getTypeName(self, i: int) -> str:
        return self._names[i] if i < len(self._names) else str(i)

    def __eq__(self, other):
        return isinstance(other, TypeSet) and self._names == other._names


class SCMigrationModel:
    """Same Inputs as the reference class (SCMigrationModel.java:38-66)."""

    def __init__(self, rateMatrix: Sequence[float], popSizes: Sequence[float],
                 typeSet: Optional[TypeSet] = None, rateMatrixScaleFactor: float = None,
                 popSizesScaleFactor: float = None, rateMatrixFlags: Sequence[bool] = None,
                 useForwardMigrationRateMatrix: bool = False, nTypes: int = None, ctx=None):
        self.rateMatrix = np.asarray(rateMatrix, np.float64).ravel()
        self.popSizes = np.asarray(popSizes, np.float64).ravel()
        self.typeSet = typeSet
        self.nTypes = typeSet.getNTypes() if typeSet is not None else (
            nTypes if nTypes is not None else len(self.popSizes))
        self.rateMatrixScaleFactor = rateMatrixScaleFactor
        self.popSizesScaleFactor = popSizesScaleFactor
        self.rateMatrixFlags = None if rateMatrixFlags is None else np.asarray(
            rateMatrixFlags, bool).ravel()
        self.useForwardMigrationRateMatrix = bool(useForwardMigrationRateMatrix)
        self._ctx = ctx
        self._unif_sym = None     # which variant is currently loaded on the device
        self.initAndValidate()

    # ---- initAndValidate (:93-144)
    def initAndValidate(self):
        d, dim = self.nTypes, len(self.rateMatrix)
        if len(self.popSizes) != d:
            raise ValueError("popSizes has %d elements for %d types" % (len(self.popSizes), d))
        if dim == d * d:
            self.rateMatrixIsSquare, self.symmetricRateMatrix = True, False
        elif dim == d * (d - 1):
            self.rateMatrixIsSquare, self.symmetricRateMatrix = False, False
        elif dim == d * (d - 1) // 2:
            self.rateMatrixIsSquare, self.symmetricRateMatrix = False, True
        else:
            raise ValueError("Migration matrix has incorrect number of elements for given "
                             "deme count.")                                   # :127-129
        if self.rateMatrixFlags is not None and len(self.rateMatrixFlags) != dim:
            raise ValueError("Migration rate flags array does not have same number of "
                             "elements as migration rate matrix.")            # :136-139
        self.dirty = True

    def getNTypes(self) -> int:
        return self.nTypes

    def getTypeSet(self):
        return self.typeSet

    # ---- getArrayOffset (:324-345)
    def getArrayOffset(self, i: int, j: int) -> int:
        if i == j:
            raise RuntimeError("Programmer error: requested migration rate array offset for "
                               "diagonal element of migration rate matrix.")
        d = self.nTypes
        if self.rateMatrixIsSquare:
            return i * d + j
        if self.symmetricRateMatrix:
            return i * (i - 1) // 2 + j if j < i else j * (j - 1) // 2 + i
        return i * (d - 1) + (j - 1 if j > i else j)

    def getRateScaleFactor(self) -> float:
        return 1.0 if self.rateMatrixScaleFactor is None else float(self.rateMatrixScaleFactor)

    def getPopSizeScaleFactor(self) -> float:
        return 1.0 if self.popSizesScaleFactor is None else float(self.popSizesScaleFactor)

    def getPopSize(self, i: int) -> float:                                   # :373-375
        return self.getPopSizeScaleFactor() * float(self.popSizes[i])

    def getBackwardRate(self, i: int, j: int) -> float:                      # :252-268
        if i == j:
            return 0.0
        if self.useForwardMigrationRateMatrix:
            return self.getForwardRate(j, i) * self.getPopSize(j) / self.getPopSize(i)
        off = self.getArrayOffset(i, j)
        if self.rateMatrixFlags is not None and not self.rateMatrixFlags[off]:
            return 0.0
        return self.getRateScaleFactor() * float(self.rateMatrix[off])

    def getForwardRate(self, i: int, j: int) -> float:                       # :285-303
        if i == j:
            return 0.0
        if self.useForwardMigrationRateMatrix:
            off = self.getArrayOffset(i, j)
            if self.rateMatrixFlags is not None and not self.rateMatrixFlags[off]:
                return 0.0
            return self.getRateScaleFactor() * float(self.rateMatrix[off])
        return self.getBackwardRate(j, i) * self.getPopSize(j) / self.getPopSize(i)

    def dense(self):
        """(m[d,d] backward rates with zero diagonal, N[d]) -- what the C-ABI receives."""
        d = self.nTypes
        m = np.zeros((d, d))
        for i in range(d):
            for j in range(d):
                if i != j:
                    m[i, j] = self.getBackwardRate(i, j)
        N = np.array([self.getPopSize(i) for i in range(d)])
        return m, N

    def setRateMatrix(self, values):
        self.rateMatrix = np.asarray(values, np.float64).ravel()
        self.dirty = True                                                    # :479-490

    def setPopSizes(self, values):
        self.popSizes = np.asarray(values, np.float64).ravel()
        self.dirty = True

    # ---- uniformization side (device): getQ / getR / getMu / getRpowN / RpowSteadyN
    def _ensure_device(self, sym: bool):
        from .uniformization import UniformizationModels
        if self.dirty or self._unif_sym != bool(sym) or getattr(self, "_um", None) is None:
            m, _ = self.dense()
            self._um = UniformizationModels(m[None, :, :], sym=bool(sym), ctx=self._ctx)
            self._unif_sym = bool(sym)
            self.dirty = False
        return self._um

    def getQ(self, symmetric: bool = False) -> np.ndarray:
        return self._ensure_device(symmetric).matrices(0)[0]

    def getR(self, symmetric: bool = False) -> np.ndarray:
        return self._ensure_device(symmetric).matrices(0)[1]

    def getMu(self, symmetric: bool = False) -> float:
        return self._ensure_device(symmetric).matrices(0)[2]

    def getRpowN(self, n: int, symmetric: bool = False) -> np.ndarray:
        pows, _ = self._ensure_device(symmetric).rpow(0, n)
        return pows[n]

    def RpowSteadyN(self, symmetric: bool = False, upto: int = 0) -> int:
        _, steady = self._ensure_device(symmetric).rpow(0, upto)
        return steady
