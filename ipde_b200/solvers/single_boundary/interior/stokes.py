"""
Single-boundary interior Stokes solver: signature mirror of ipde/solvers/single_boundary/interior/stokes.py:29-164
-- `StokesSolver(ebdy, bump, bump_loc, solver_type, ASS)(fu, fv, fur, fvr, **kwargs) -> (uc, vc, pc, ur, vr, pr)`
-- served by the multi-boundary device solver on a one-boundary collection.
"""
from ._adapter import collection_of, pack, unpack
from ...multi_boundary.stokes import StokesSolver as _Multi


class StokesSolver(object):
    def __init__(self, ebdy, bump, bump_loc, solver_type='spectral', ASS=None):
        self.ebdyc = collection_of(ebdy)
        self.ebdy = self.ebdyc[0]
        self.solver_type = solver_type
        self.interior = self.ebdy.interior
        if not self.ebdyc.bumpy_readied:
            self.ebdyc.ready_bump(bump, bump_loc, self.ebdy.radial_width)
        self._solver = _Multi(self.ebdyc, solver_type=solver_type)
        self.annular_solver = self._solver.helpers[0].annular_solver

    def __call__(self, fu, fv, fur, fvr, **kwargs):
        u, v, p = self._solver(pack(self.ebdyc, fu, fur), pack(self.ebdyc, fv, fvr), **kwargs)
        (uc, ur), (vc, vr), (pc, pr) = unpack(u), unpack(v), unpack(p)
        return uc, vc, pc, ur, vr, pr
