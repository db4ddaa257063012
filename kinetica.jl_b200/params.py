"""`ODESimulationParams` — fields, defaults and validation of reference src/solving/params.jl:3-110."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional, Tuple, Union


class B200Rodas4:
    """Marker usable as `pars.solver`: the batched Rodas4 integrator of libkinetica_b200."""


@dataclass
class ODESimulationParams:
    tspan: Tuple[float, float]
    u0: Any                        # Dict[str, float] or vector
    solver: Any = None
    jac: bool = True
    sparse: bool = True
    abstol: float = 1.0e-10
    reltol: float = 1.0e-8
    adaptive_tols: bool = True
    update_tols: bool = False
    solve_chunks: bool = True
    solve_chunkstep: float = 1e-3
    maxiters: int = 100000
    ban_negatives: bool = False
    progress: bool = False
    save_interval: Optional[float] = None
    low_k_cutoff: Union[float, str] = "auto"
    low_k_maxconc: float = 2.0
    allow_short_u0: bool = False

    def __post_init__(self):
        self.tspan = (float(self.tspan[0]), float(self.tspan[1]))
        if self.tspan[0] >= self.tspan[1]:
            raise ValueError(f"Invalid time span: Start = {self.tspan[0]}, End = {self.tspan[1]}")
        if isinstance(self.low_k_cutoff, str):
            if self.low_k_cutoff not in ("auto", "none"):
                raise ValueError("low_k_cutoff must be a numerical value or one of [:auto, :none]")
        elif self.low_k_cutoff < 0:
            raise ValueError("low_k_cutoff must be a positive number or one of [:auto, :none]")
        if self.solve_chunks:
            q = self.tspan[1] / self.solve_chunkstep
            if q != int(q):       # Int(tspan[2]/chunkstep) InexactError, params.jl:87-97
                raise ValueError("Simulation timespan is not divisible by requested chunkwise simulation step size")
        if self.solve_chunks and self.save_interval is not None and self.save_interval > self.solve_chunkstep:
            raise ValueError("Solution save interval must be less than chunkwise simulation step size")
        if self.solver is None:
            self.solver = B200Rodas4()
