"""PoissonSolver -- mirror of ipde/solvers/multi_boundary/poisson.py:22-65."""
from .scalar import ScalarSolver
from ..internals.poisson import PoissonHelper


class PoissonSolver(ScalarSolver):
    kernel_kind = 0
    grid_kind = 0
    k = 0.0

    def __init__(self, ebdyc, solver_type='spectral', AS_list=None, grid_backend=None, shard_boundaries=False):
        super().__init__(ebdyc, solver_type, AS_list, grid_backend, shard_boundaries=shard_boundaries)

    def _get_helper(self, ebdy, helper, owned=True):
        return PoissonHelper(ebdy, helper, grid_backend=self.grid_backend, owned=owned)
