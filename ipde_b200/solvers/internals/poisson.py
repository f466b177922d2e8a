"""PoissonHelper -- mirror of ipde/solvers/internals/poisson.py:14-36."""
from .scalar import ScalarHelper
from ...annular import AnnularPoissonSolver
from ...qfs import Laplace_QFS


class PoissonHelper(ScalarHelper):
    kernel_kind = 0
    k = 0.0

    def __init__(self, ebdy, annular_solver=None, grid_backend='pybie2d', owned=True):
        super().__init__(ebdy, annular_solver, grid_backend, owned=owned)

    def _define_annular_solver(self):
        self.annular_solver = AnnularPoissonSolver(self.AAG)

    def _get_qfs(self):
        tol = getattr(self.ebdy, 'qfs_tolerance', 1e-12)
        self.interface_qfs_g = Laplace_QFS(self.ebdy.interface, self.interior, True, True, tol=min(tol, 1e-12), u2s=False, operators=self.owned)
        self.interface_qfs_r = Laplace_QFS(self.ebdy.interface, not self.interior, True, True, tol=min(tol, 1e-12),
                                           operators=self.owned)
