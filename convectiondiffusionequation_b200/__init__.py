"""B200-native (E)HDG convection-diffusion hot path (drop-in for the NGSolve-scripted
Convection_Diffusion._solveEHDG of aqibrahimbt/convectiondiffusionequation)."""
from .coefficients import LayerExpr, SymExpr, boundary_layer, x, y          # noqa: F401
from .mesh import TriMesh, mesh_for_size, unit_square_mesh          # noqa: F401

__all__ = ["LayerExpr", "SymExpr", "boundary_layer", "x", "y", "TriMesh", "mesh_for_size", "unit_square_mesh"]
