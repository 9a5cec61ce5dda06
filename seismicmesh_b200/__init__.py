"""seismicmesh_b200 -- B200-native (sm_100a CUDA) implementation of the DistMesh force-iteration
loop of SeismicMesh's ``generate_mesh`` / ``sliver_removal``, behind the reference's Python API.

Drop-in for that path only: ``generate_mesh``, ``sliver_removal``, the SDF primitives and
combinators, and ``SizeFunction``.  Importing this package loads ``libdistmesh_b200.so``; there
is no CPU fallback.
"""
from . import decomp, geometry, migration, sizing
from .generation import generate_mesh, last_run_stats, sliver_removal
from .geometry import (
    Ball,
    Cube,
    Cylinder,
    Difference,
    Disk,
    Intersection,
    Prism,
    Rectangle,
    Repeat,
    Torus,
    Union,
)
from .sizing import GridInterpolant, SizeFunction, get_sizing_function_from_segy, read_velocity_model

__version__ = "0.1.0"

__all__ = [
    "__version__",
    "geometry",
    "sizing",
    "decomp",
    "migration",
    "read_velocity_model",
    "Rectangle",
    "Cube",
    "Cylinder",
    "Disk",
    "Union",
    "Torus",
    "Prism",
    "Ball",
    "Intersection",
    "Difference",
    "Repeat",
    "generate_mesh",
    "sliver_removal",
    "SizeFunction",
    "get_sizing_function_from_segy",
    "GridInterpolant",
    "last_run_stats",
]
