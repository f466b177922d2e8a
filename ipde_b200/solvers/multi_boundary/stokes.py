"""StokesSolver -- mirror of ipde/solvers/multi_boundary/stokes.py:6-50 (mu = 1)."""
from .vector import VectorSolver
from ..internals.stokes import StokesHelper


class StokesSolver(VectorSolver):
    def __init__(self, ebdyc, solver_type='spectral', helpers=None, grid_backend=None, shard_boundaries=False):
        if not ebdyc.bumpy_readied:
            raise Exception('Stokes solver requires embedded boundary collection with a bump function.')
        super().__init__(ebdyc, solver_type, helpers, grid_backend=grid_backend, shard_boundaries=shard_boundaries)

    def _get_helper_compatability(self, ebdy, helper):
        """0: start over, 1: annular solver reusable, 2: helper reusable as is"""
        if helper is None:
            return 0
        if ebdy.bdy.N != helper.ebdy.bdy.N:
            return 0
        return 1 if helper.ebdy is not ebdy else 2

    def _get_helper(self, ebdy, helper):
        c = self._get_helper_compatability(ebdy, helper)
        if c == 0:
            return StokesHelper(ebdy)
        if c == 1:
            return StokesHelper(ebdy, helper.annular_solver)
        return helper
