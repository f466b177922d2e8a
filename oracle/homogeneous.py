"""
ORACLE (test infrastructure, CPU, NumPy/SciPy) -- never imported by the product.

CPU statement of the homogeneous Dirichlet correction of ipde_b200/solvers/multi_boundary/homogeneous.py: the step
the reference's callers perform after the inhomogeneous solve
(examples/multi_modified_helmholtz.py:113-177, examples/interior_modified_helmholtz.py:73-86).

The reference's route (dense second-kind BIE from pybie2d self-quadratures + numpy.linalg.solve + QFS) cannot run
here: pybie2d and qfs are absent, so this step has no golden vectors from the reference -- parity
unpinned.  What pins it instead: (i) uniqueness of the homogeneous Dirichlet solution -- after the correction the
manufactured solution of the examples must be reproduced (tests/test_oracle.py), (ii) this all-pairs NumPy statement
against the device implementation on identical inputs (tests/test_gpu_solvers.py).

    tau_i + sum_{j != i} S(bdy_i <- src_j) u2s_j(tau_j) = bc_i - u|_i ,   u += sum_j S(. <- src_j) u2s_j(tau_j)
"""
import numpy as np
import scipy.sparse.linalg
import scipy.special

from .solvers import layer_apply


class OracleDirichletCorrection(object):
    def __init__(self, ebdyc, kind, k, qfs_list):
        """kind: 'poisson' | 'modified_helmholtz'; qfs_list: one ipde_b200.qfs object per boundary (host built, with
        the u2s operators) on ebdy.bdy, sources outside the physical domain"""
        self.ebdyc, self.kind, self.k, self.qfs = ebdyc, kind, float(k), qfs_list
        self.boff = np.concatenate([[0], np.cumsum([e.bdy.N for e in ebdyc])])
        self.iterations = 0

    def _weighted_sources(self, tau):
        return [q.u2s(tau[self.boff[i]:self.boff[i + 1]])*q.source.weights for i, q in enumerate(self.qfs)]

    def _apply(self, tau):
        sig = self._weighted_sources(tau)
        out = np.array(tau, dtype=float)
        for i, ebdy in enumerate(self.ebdyc):
            for j, q in enumerate(self.qfs):
                if i != j:
                    out[self.boff[i]:self.boff[i + 1]] += layer_apply(self.kind, self.k, q.source.x, q.source.y, sig[j],
                                                                      ebdy.bdy.x, ebdy.bdy.y)
        return out

    def __call__(self, ue, bc, tol=1e-13, restart=50, maxiter=10):
        e = self.ebdyc
        bv = np.concatenate([ebdy.interpolate_radial_to_boundary(ue[i]) for i, ebdy in enumerate(e)])
        g = np.asarray(bc, dtype=float).ravel() - bv
        if len(self.qfs) == 1:
            tau, self.iterations = g, 0
        else:
            count = [0]
            def cb(_):
                count[0] += 1
            A = scipy.sparse.linalg.LinearOperator((g.size, g.size), matvec=self._apply, dtype=float)
            tau, info = scipy.sparse.linalg.gmres(A, g, rtol=tol, atol=0.0, restart=restart, maxiter=maxiter,
                                                  callback=cb, callback_type='pr_norm')
            self.iterations = count[0]
            if info != 0:
                raise Exception('oracle Dirichlet correction: GMRES did not converge')
        sig = self._weighted_sources(tau)
        pts = e.grid_and_radial_pts
        u = np.array(np.asarray(ue), dtype=float)
        for q, s in zip(self.qfs, sig):
            u += layer_apply(self.kind, self.k, q.source.x, q.source.y, s, pts.x, pts.y)
        return u


class OracleNeumannCorrection(object):
    """
    CPU statement of NeumannCorrection (examples/multi_modified_helmholtz_neumann.py:97-163; modified Helmholtz):
        (A tau)_i = d/dn_i sum_j S(bdy_i <- src_j) QFS_j(tau_j) = bc_n,i - d/dn u|_i ,  u += sum_j S(. <- src_j) QFS_j(tau_j)
    all pairs in NumPy, K1 from SciPy.  Parity unpinned against the reference (pybie2d's Kress self-quadrature is
    absent); pinned by the analytic solution and against the device path on identical inputs.
    """
    def __init__(self, ebdyc, k, qfs_list):
        self.ebdyc, self.k, self.qfs = ebdyc, float(k), qfs_list
        self.boff = np.concatenate([[0], np.cumsum([e.bdy.N for e in ebdyc])])
        self.iterations = 0

    def _weighted_sources(self, tau):
        return [q([tau[self.boff[i]:self.boff[i + 1]]])*q.source.weights for i, q in enumerate(self.qfs)]

    def _dn(self, sx, sy, qw, b):
        dx, dy = b.x[:, None] - sx[None, :], b.y[:, None] - sy[None, :]
        r = np.hypot(dx, dy)
        dn = dx*b.normal_x[:, None] + dy*b.normal_y[:, None]
        return (-self.k*scipy.special.k1(self.k*r)*dn/r/(2*np.pi)).dot(qw)

    def _apply(self, tau):
        sig = self._weighted_sources(tau)
        out = np.zeros(self.boff[-1])
        for i, ebdy in enumerate(self.ebdyc):
            for j, q in enumerate(self.qfs):
                out[self.boff[i]:self.boff[i + 1]] += self._dn(q.source.x, q.source.y, sig[j], ebdy.bdy)
        return out

    def __call__(self, ue, bc_n, tol=1e-13, restart=80, maxiter=10):
        e = self.ebdyc
        bn = np.concatenate([ebdy.interpolate_radial_to_boundary_normal_derivative(ue[i]) for i, ebdy in enumerate(e)])
        g = np.asarray(bc_n, dtype=float).ravel() - bn
        count = [0]
        def cb(_):
            count[0] += 1
        A = scipy.sparse.linalg.LinearOperator((g.size, g.size), matvec=self._apply, dtype=float)
        tau, info = scipy.sparse.linalg.gmres(A, g, rtol=tol, atol=0.0, restart=restart, maxiter=maxiter, callback=cb,
                                              callback_type='pr_norm')
        self.iterations = count[0]
        if info != 0:
            raise Exception('oracle Neumann correction: GMRES did not converge')
        sig = self._weighted_sources(tau)
        pts = e.grid_and_radial_pts
        u = np.array(np.asarray(ue), dtype=float)
        for q, s in zip(self.qfs, sig):
            u += layer_apply('modified_helmholtz', self.k, q.source.x, q.source.y, s, pts.x, pts.y)
        return u
