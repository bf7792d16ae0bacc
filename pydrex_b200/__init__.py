"""pydrex_b200: the CPO-integration hot path of PyDRex as hand-written sm_100a CUDA.

Drop-in for `pydrex.core.derivatives`, `pydrex.minerals.Mineral.update_orientations` /
`update_all` / `voigt_averages` and `pydrex.diagnostics.misorientation_index(es)`, plus the
batched `ParticleEnsemble` fast path.  See DESIGN.md and INTEGRATION.md.
"""

from . import _lib, geometry, pathlines, stats, tensors, utils, velocity
from .core import (
    DefaultParams,
    DeformationRegime,
    MineralFabric,
    MineralPhase,
    derivatives,
    get_crss,
)
from .diagnostics import (
    bingham_average,
    elasticity_components,
    misorientation_index,
    misorientation_indices,
    symmetry_pgr,
)
from .ensemble import ParticleEnsemble, gather_to_rank0, shard_range
from .exceptions import Error, IterationError
from .geometry import LatticeSystem, symmetry_operations
from .minerals import (
    OLIVINE_PRIMARY_AXIS,
    OLIVINE_SLIP_SYSTEMS,
    Mineral,
    StiffnessTensors,
    update_all,
    voigt_averages,
)
from .pathlines import get_pathline, get_pathlines
from .stats import misorientation_hist, misorientations_random, resample_orientations

__all__ = [
    "DefaultParams", "DeformationRegime", "MineralFabric", "MineralPhase", "derivatives",
    "get_crss", "bingham_average", "elasticity_components", "symmetry_pgr", "misorientation_index", "misorientation_indices", "ParticleEnsemble",
    "gather_to_rank0", "shard_range", "Error", "IterationError", "LatticeSystem",
    "symmetry_operations", "Mineral", "StiffnessTensors", "update_all", "voigt_averages",
    "misorientation_hist", "misorientations_random", "resample_orientations", "get_pathline", "get_pathlines", "velocity", "pathlines", "tensors", "utils", "OLIVINE_PRIMARY_AXIS",
    "OLIVINE_SLIP_SYSTEMS",
]
