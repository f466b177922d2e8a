"""ModifiedHelmholtzHelper -- mirror of ipde/solvers/internals/modified_helmholtz.py:13-37."""
from .scalar import ScalarHelper
from ...annular import AnnularModifiedHelmholtzSolver
from ...qfs import Modified_Helmholtz_QFS


class ModifiedHelmholtzHelper(ScalarHelper):
    kernel_kind = 1

    def __init__(self, ebdy, annular_solver=None, k=1.0, source_upsample_factor=1.0, grid_backend='pybie2d', owned=True):
        self.k = k
        self.source_upsample_factor = source_upsample_factor
        super().__init__(ebdy, annular_solver, grid_backend, owned=owned)

    def _define_annular_solver(self):
        self.annular_solver = AnnularModifiedHelmholtzSolver(self.AAG, k=self.k)

    def _get_qfs(self):
        tol = min(getattr(self.ebdy, 'qfs_tolerance', 1e-12), 1e-12)
        suf = self.source_upsample_factor
        self.interface_qfs_g = Modified_Helmholtz_QFS(self.ebdy.interface, self.interior, True, True, self.k,
                                                      source_upsample_factor=suf, tol=tol, u2s=False, operators=self.owned)
        self.interface_qfs_r = Modified_Helmholtz_QFS(self.ebdy.interface, not self.interior, True, True, self.k,
                                                      source_upsample_factor=suf, tol=tol, operators=self.owned)
