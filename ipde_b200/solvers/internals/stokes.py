"""StokesHelper -- mirror of ipde/solvers/internals/stokes.py:10-35 (mu = 1)."""
import numpy as np
from .vector import VectorHelper
from ...annular import AnnularStokesSolver
from ...qfs import Stokes_QFS
from ...device import to_dev, layer_apply_stokes


class StokesHelper(VectorHelper):
    def __init__(self, ebdy, annular_solver=None):
        super().__init__(ebdy, annular_solver)

    def _define_annular_solver(self):
        self.annular_solver = AnnularStokesSolver(self.AAG, mu=1.0)

    def _get_qfs(self):
        tol = min(getattr(self.ebdy, 'qfs_tolerance', 1e-12), 1e-12)
        self.interface_qfs_g = Stokes_QFS(self.ebdy.interface, self.interior, True, True, tol=tol, u2s=False)
        self.interface_qfs_r = Stokes_QFS(self.ebdy.interface, not self.interior, True, True, tol=tol)

    def _define_layer_apply(self):
        def Layer_Apply(src, trg, f):
            """(u, v, p) at trg of stokeslets f (2, N_src) at src -- the reference's SFMM call (stokes.py:26-35)"""
            f = np.asarray(f, dtype=float).reshape(2, -1)
            q0, q1 = to_dev(f[0]*src.weights), to_dev(f[1]*src.weights)
            out = layer_apply_stokes(to_dev(src.x), to_dev(src.y), q0, q1, to_dev(trg.x), to_dev(trg.y)).cpu().numpy()
            return out[0], out[1], out[2]
        self.Layer_Apply = Layer_Apply
