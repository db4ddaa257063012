"""kinetica_b200 — B200-native kinetic-solve hot path of Kinetica.jl behind the reference's
calculator / ODESimulationParams / ConditionSet API.  Sources live in `kinetica.jl_b200/`;
import as `kinetica_b200`."""
from .calculator import (AbstractKineticCalculator, CollisionTheoryCalculator, DummyKineticCalculator,
                         EyringCalculator, PrecalculatedArrheniusCalculator)
from .conditions import (ConditionSet, DoubleRampGradientProfile, LinearDirectProfile,
                         LinearGradientProfile, NullDirectProfile, NullGradientProfile,
                         StaticConditionProfile, create_savepoints, tconvert)
from .io import load_output, save_output
from .network import RxData, SpeciesData
from .parallel import solve_network_sharded
from .params import B200Rodas4, ODESimulationParams
from .seeds import identify_next_seeds, identify_next_seeds_ensemble
from .solve import (B200EnsembleODESolve, EnsembleSolver, ODESolveOutput, RxFilter, StaticODESolve,
                    VariableODESolve, solve_network)
