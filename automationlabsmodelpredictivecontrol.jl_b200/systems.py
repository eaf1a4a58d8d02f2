"""Minimal stand-ins for the MathematicalSystems.jl / LazySets.jl objects the reference's design code reads
(`system.A .B .X .U .f .statedim .inputdim`, `LazySets.vertices_list`), so the Python host mirror can be driven the
way the reference's tests drive the Julia package (test/computation_mpc_test.jl:981-1012)."""
from __future__ import annotations

import dataclasses
from typing import Any

import numpy as np


@dataclasses.dataclass
class Hyperrectangle:
    low: np.ndarray
    high: np.ndarray

    def __post_init__(self):
        self.low = np.asarray(self.low, float); self.high = np.asarray(self.high, float)
        if self.low.shape != self.high.shape or np.any(self.low > self.high):
            raise ValueError("Hyperrectangle: need low <= high of equal length")


@dataclasses.dataclass
class ConstrainedLinearControlDiscreteSystem:
    A: np.ndarray
    B: np.ndarray
    X: Hyperrectangle
    U: Hyperrectangle

    def __post_init__(self):
        self.A = np.asarray(self.A, float); self.B = np.asarray(self.B, float)

    @property
    def statedim(self): return self.A.shape[0]
    @property
    def inputdim(self): return self.B.shape[1]


@dataclasses.dataclass
class ConstrainedBlackBoxControlDiscreteSystem:
    f: Any                     # a Chain (nn.py) -- the Flux model of the reference
    statedim: int
    inputdim: int
    X: Hyperrectangle
    U: Hyperrectangle
