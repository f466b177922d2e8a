"""ModifiedHelmholtzSolver -- mirror of ipde/solvers/multi_boundary/modified_helmholtz.py:8-67."""
from .scalar import ScalarSolver
from ..internals.modified_helmholtz import ModifiedHelmholtzHelper


class ModifiedHelmholtzSolver(ScalarSolver):
    kernel_kind = 1
    grid_kind = 1

    def __init__(self, ebdyc, k, solver_type='spectral', helpers=None, grid_backend='pybie2d',
                 source_upsample_factor=1.0, shard_boundaries=False):
        self.k = k
        self.source_upsample_factor = source_upsample_factor
        super().__init__(ebdyc, solver_type, helpers, grid_backend, shard_boundaries=shard_boundaries)

    def _get_helper_combatibility(self, ebdy, helper):
        """0: start over, 1: annular solver reusable, 2: helper reusable as is"""
        if helper is None:
            return 0
        if self.k != helper.k or ebdy.bdy.N != helper.ebdy.bdy.N:
            return 0
        if self.source_upsample_factor != helper.source_upsample_factor:
            return 0
        return 1 if helper.ebdy is not ebdy else 2

    def _get_helper(self, ebdy, helper, owned=True):
        c = self._get_helper_combatibility(ebdy, helper)
        if c == 0:
            return ModifiedHelmholtzHelper(ebdy, k=self.k, source_upsample_factor=self.source_upsample_factor,
                                           grid_backend=self.grid_backend, owned=owned)
        if c == 1:
            # a donor helper built with owned=False (boundary sharding) has no annular solver to lend
            donor = getattr(helper, 'annular_solver', None)
            return ModifiedHelmholtzHelper(ebdy, donor, k=self.k, source_upsample_factor=self.source_upsample_factor,
                                           grid_backend=self.grid_backend, owned=owned)
        if owned and not helper.owned:
            return ModifiedHelmholtzHelper(ebdy, k=self.k, source_upsample_factor=self.source_upsample_factor,
                                           grid_backend=self.grid_backend, owned=True)
        return helper
