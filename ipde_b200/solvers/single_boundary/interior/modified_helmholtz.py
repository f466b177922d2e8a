"""
Single-boundary interior modified-Helmholtz solver: signature mirror of
ipde/solvers/single_boundary/interior/modified_helmholtz.py:15-102 -- `ModifiedHelmholtzSolver(ebdy, k, solver_type,
AMHS)(f, fr, **kwargs) -> (uc, ur)` with `f` the (Nx, Ny) grid forcing and `fr` the (M, n) radial forcing --
served by the multi-boundary device solver on a one-boundary collection (the reference file itself targets an
EmbeddedBoundary API that no longer exists).
"""
from ._adapter import collection_of, pack, unpack
from ...multi_boundary.modified_helmholtz import ModifiedHelmholtzSolver as _Multi


class ModifiedHelmholtzSolver(object):
    def __init__(self, ebdy, k, solver_type='spectral', AMHS=None):
        self.ebdyc = collection_of(ebdy)
        self.ebdy = self.ebdyc[0]
        self.k, self.solver_type = k, solver_type
        self.interior = self.ebdy.interior
        self._solver = _Multi(self.ebdyc, k, solver_type=solver_type, helpers=None if AMHS is None else [AMHS])
        self.annular_solver = self._solver.helpers[0].annular_solver

    def get_bv(self, ur):
        return self._solver.helpers[0].get_boundary_values(ur)

    def get_bn(self, ur):
        return self._solver.helpers[0].get_boundary_normal_derivatives(ur)

    def __call__(self, f, fr, **kwargs):
        return unpack(self._solver(pack(self.ebdyc, f, fr), **kwargs))
