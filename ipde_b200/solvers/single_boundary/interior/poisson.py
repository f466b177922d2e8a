"""
Single-boundary interior Poisson solver: signature mirror of ipde/solvers/single_boundary/interior/poisson.py:20-111
-- `PoissonSolver(ebdy, bump, bump_loc, solver_type, APS)(f, fr, **kwargs) -> (uc, ur)` -- served by the
multi-boundary device solver on a one-boundary collection.
"""
from ._adapter import collection_of, pack, unpack
from ...multi_boundary.poisson import PoissonSolver as _Multi


class PoissonSolver(object):
    def __init__(self, ebdy, bump, bump_loc, solver_type='spectral', APS=None):
        self.ebdyc = collection_of(ebdy)
        self.ebdy = self.ebdyc[0]
        self.solver_type = solver_type
        self.interior = self.ebdy.interior
        if not self.ebdyc.bumpy_readied:
            self.ebdyc.ready_bump(bump, bump_loc, self.ebdy.radial_width)
        self._solver = _Multi(self.ebdyc, solver_type=solver_type, AS_list=None if APS is None else [APS])
        self.annular_solver = self._solver.helpers[0].annular_solver

    def get_bv(self, ur):
        return self._solver.helpers[0].get_boundary_values(ur)

    def get_bn(self, ur):
        return self._solver.helpers[0].get_boundary_normal_derivatives(ur)

    def __call__(self, f, fr, **kwargs):
        return unpack(self._solver(pack(self.ebdyc, f, fr), **kwargs))
