"""
Homogeneous (boundary-condition) correction -- the step every caller of the reference performs right after the
inhomogeneous solve:

    examples/interior_modified_helmholtz.py:73-86       single boundary, Dirichlet        DirichletCorrection
    examples/multi_modified_helmholtz.py:113-177        boundary collection, Dirichlet    DirichletCorrection
    examples/poisson_for_paper.py:88-105                Poisson, Dirichlet                DirichletCorrection
    examples/multi_modified_helmholtz_neumann.py:97-163 boundary collection, Neumann      NeumannCorrection
    examples/interior_stokes_with_multi.py:98-125       Stokes, prescribed velocity       StokesDirichletCorrection

The reference solves a dense second-kind integral equation for a double-layer density on all boundaries (pybie2d
self-quadratures + numpy.linalg.solve), turns it into effective sources with QFS and sums them onto
`ebdyc.grid_and_radial_pts`.  Here the homogeneous solution is represented by the same kind of effective sources on
the QFS source curve of every boundary (outside the physical domain), found from the boundary data directly:

    tau_i  : values the homogeneous solution has to take on boundary i           (unknown blocks)
    sigma_i = u2s_i(tau_i)                                                       (QFS boundary -> source operator)
    tau_i + sum_{j != i} S(bdy_i <- src_j) sigma_j = bc_i - u_inhomogeneous|_i   (well conditioned: I + compact)

one boundary: tau = g, no solve; several: GMRES on the block system (the Krylov bookkeeping runs on the host on
vectors of sum n_b numbers, every operator application -- QFS mat-vecs, boundary <- source sums -- on the device).
The sources are then summed onto all physical grid points and all radial points by the same device evaluators the
inhomogeneous solve uses (spectral Ewald, or all pairs for small problems), and added to `ue`.
Both routes compute the unique solution of the homogeneous Dirichlet problem; they agree to the QFS tolerance.
"""
import numpy as np
import scipy.sparse.linalg

from ...device import to_dev, empty, layer_apply_scalar, layer_apply_scalar_dn, FreeSpaceEvaluator
from ...embedded_function import EmbeddedFunction
from ...qfs import Laplace_QFS, Modified_Helmholtz_QFS, Stokes_QFS
from ..internals.scalar import DeviceQFS, device_qfs


class DirichletCorrection(object):
    """
    corr = DirichletCorrection(solver)            # solver: PoissonSolver / ModifiedHelmholtzSolver of ipde_b200
    ue = solver(f, **gmres_kwargs)
    u = corr(ue, bc)                              # bc: BoundaryFunction / concatenated boundary values
    """
    EWALD_THRESHOLD = 2.0e8         # sources x targets above which the spectral-Ewald evaluator is used

    def __init__(self, solver, tol=1e-12, evaluation='auto'):
        self.solver = solver
        e = self.ebdyc = solver.ebdyc
        self.kind, self.k = solver.kernel_kind, float(solver.k)
        self.qfs, self.dqfs = [], []
        for ebdy in e:
            q = self._build_qfs(ebdy, tol)
            self.qfs.append(q)
            self.dqfs.append(device_qfs(q))
        self.nb = [ebdy.bdy.N for ebdy in e]
        self.boff = np.concatenate([[0], np.cumsum(self.nb)])
        self.d_bx = [to_dev(ebdy.bdy.x) for ebdy in e]
        self.d_by = [to_dev(ebdy.bdy.y) for ebdy in e]
        sx = np.concatenate([q.source.x for q in self.qfs])
        sy = np.concatenate([q.source.y for q in self.qfs])
        self.soff = np.concatenate([[0], np.cumsum([q.source.N for q in self.qfs])])
        self.d_sx, self.d_sy = to_dev(sx), to_dev(sy)
        pts = e.grid_and_radial_pts
        g = e.grid
        if evaluation == 'auto':
            evaluation = 'ewald' if float(sx.size)*float(pts.N) > self.EWALD_THRESHOLD else 'direct'
        if evaluation == 'ewald' and abs(g.xh - g.yh) > 1e-14*g.xh:
            evaluation = 'direct'
        self.evaluation = evaluation
        self.n_phys = int(e.phys.sum())
        if evaluation == 'ewald':
            pix, piy = np.nonzero(e.phys)
            rp = e.radial_pts
            self.evaluator = FreeSpaceEvaluator(self.kind, self.k, float(g.xv[0]), float(g.yv[0]), float(g.xh), sx, sy,
                                                grid_nodes=(pix, piy), grid_out_pos=e.phys_inds[e.phys],
                                                points=(rp.x, rp.y))
        else:
            self.d_tx, self.d_ty = to_dev(pts.x), to_dev(pts.y)
        self.iterations = 0

    def _build_qfs(self, ebdy, tol):
        """boundary values -> sources (u2s) only"""
        if self.kind == 0:
            return Laplace_QFS(ebdy.bdy, ebdy.interior, True, False, tol=tol, b2c=False)
        return Modified_Helmholtz_QFS(ebdy.bdy, ebdy.interior, True, False, self.k, tol=tol, b2c=False)

    # ---- boundary data -> weighted source strengths of every boundary (one flat device vector)
    def _sources(self, tau_blocks, out):
        for i, dq in enumerate(self.dqfs):
            s = dq.u2s(tau_blocks[i])
            out[self.soff[i]:self.soff[i + 1]] = s*dq.sw
        return out

    def _apply(self, tau_host):
        """(A tau)_i = tau_i + sum_{j != i} S(bdy_i <- src_j) u2s_j(tau_j)"""
        n = len(self.dqfs)
        t = to_dev(np.ascontiguousarray(tau_host, dtype=float))
        blocks = [t[self.boff[i]:self.boff[i + 1]] for i in range(n)]
        q = self._sources(blocks, empty((int(self.soff[-1]),)))
        res = t.clone()
        for i in range(n):
            for j in range(n):
                if i != j:
                    layer_apply_scalar(self.kind, self.k, self.d_sx[self.soff[j]:self.soff[j + 1]],
                                       self.d_sy[self.soff[j]:self.soff[j + 1]], q[self.soff[j]:self.soff[j + 1]],
                                       self.d_bx[i], self.d_by[i], out=res[self.boff[i]:self.boff[i + 1]], accumulate=True)
        return res.cpu().numpy()

    def densities(self, g, tol=1e-13, restart=50, maxiter=10):
        """boundary data tau (host vector, all boundaries) with the homogeneous solution equal to g on every boundary"""
        g = np.asarray(g, dtype=float).ravel()
        if len(self.dqfs) == 1:
            self.iterations = 0
            return g
        count = [0]
        def cb(_):
            count[0] += 1
        A = scipy.sparse.linalg.LinearOperator((g.size, g.size), matvec=self._apply, dtype=float)
        tau, info = scipy.sparse.linalg.gmres(A, g, rtol=tol, atol=0.0, restart=restart, maxiter=maxiter, callback=cb,
                                              callback_type='pr_norm')
        self.iterations = count[0]
        if info != 0:
            raise Exception('DirichletCorrection: GMRES on the boundary system did not converge (info = %d)' % info)
        return tau

    def __call__(self, ue, bc, **kwargs):
        bv = self.solver.get_boundary_values(ue)
        g = np.asarray(bc, dtype=float).ravel() - np.asarray(bv, dtype=float).ravel()
        tau = self.densities(g, **kwargs)
        return self._add_field(ue, tau)

    def _add_field(self, ue, tau):
        """ue + field of the sources of the boundary densities tau on all physical grid points and radial points"""
        t = to_dev(tau)
        q = self._sources([t[self.boff[i]:self.boff[i + 1]] for i in range(len(self.dqfs))],
                          empty((int(self.soff[-1]),)))
        u = to_dev(np.ascontiguousarray(np.asarray(ue), dtype=float))
        if self.evaluation == 'ewald':
            self.evaluator(q, grid_out=u, point_out=u[self.n_phys:], accumulate_points=True)
        else:
            layer_apply_scalar(self.kind, self.k, self.d_sx, self.d_sy, q, self.d_tx, self.d_ty, out=u, accumulate=True)
        return EmbeddedFunction(self.ebdyc, linear_data=u.cpu().numpy())


class NeumannCorrection(DirichletCorrection):
    """
    Homogeneous correction for NEUMANN data (examples/multi_modified_helmholtz_neumann.py:97-163): after
    `ue = solver(f)`, find the homogeneous solution whose normal derivative on every boundary is
    bc_n - d/dn ue, and add it on all grid and radial points.

    The reference solves the second-kind equation (S' -+ I/2) tau = g for a single-layer density with pybie2d's Kress
    self-quadrature (its `s_singular` / `s_only` blocks, :117-131), then tau -> QFS sources -> Layer_Apply (:141-158).
    Here the same density is the unknown, but the operator is applied THROUGH the sources:

        sigma_j = QFS_j(tau_j)                        density -> effective sources (single-layer QFS, as :146)
        (A tau)_i = d/dn_i sum_j S(bdy_i <- src_j) sigma_j     k_layer_scalar_dn, every pair (i, j), i = j included

    Since the sources of QFS_j carry the single-layer potential of tau_j on the physical side, A is the reference's
    second-kind operator (GMRES: a few tens of iterations) without a singular quadrature; and because the correction
    that is finally added IS the field of these sources -- an exact solution of the homogeneous equation whose normal
    derivative is collocated at the boundary nodes -- its accuracy is that of the source representation, not of the
    quadrature.  Modified Helmholtz only (the pure Neumann Poisson problem is singular; the reference has no such
    example).  Normals: `bdy.normal_x/y`, the convention of `solver.get_boundary_normal_derivatives`.
    """

    def __init__(self, solver, tol=1e-12, evaluation='auto'):
        if solver.kernel_kind == 0:
            raise Exception('NeumannCorrection: modified Helmholtz only (the pure Neumann Poisson problem is singular)')
        DirichletCorrection.__init__(self, solver, tol=tol, evaluation=evaluation)
        self.d_bnx = [to_dev(ebdy.bdy.normal_x) for ebdy in self.ebdyc]
        self.d_bny = [to_dev(ebdy.bdy.normal_y) for ebdy in self.ebdyc]

    def _build_qfs(self, ebdy, tol):
        """single-layer density -> sources (examples/multi_modified_helmholtz_neumann.py:146)"""
        return Modified_Helmholtz_QFS(ebdy.bdy, ebdy.interior, True, False, self.k, tol=tol, u2s=False)

    def _sources(self, tau_blocks, out):
        for i, dq in enumerate(self.dqfs):
            s = dq(tau_blocks[i])
            out[self.soff[i]:self.soff[i + 1]] = s*dq.sw
        return out

    def _apply(self, tau_host):
        n = len(self.dqfs)
        t = to_dev(np.ascontiguousarray(tau_host, dtype=float))
        blocks = [t[self.boff[i]:self.boff[i + 1]] for i in range(n)]
        q = self._sources(blocks, empty((int(self.soff[-1]),)))
        res = empty((int(self.boff[-1]),))
        for i in range(n):
            layer_apply_scalar_dn(self.kind, self.k, self.d_sx, self.d_sy, q, self.d_bx[i], self.d_by[i], self.d_bnx[i],
                                  self.d_bny[i], out=res[self.boff[i]:self.boff[i + 1]])
        return res.cpu().numpy()

    def densities(self, g, tol=1e-13, restart=80, maxiter=10):
        g = np.asarray(g, dtype=float).ravel()
        count = [0]
        def cb(_):
            count[0] += 1
        A = scipy.sparse.linalg.LinearOperator((g.size, g.size), matvec=self._apply, dtype=float)
        tau, info = scipy.sparse.linalg.gmres(A, g, rtol=tol, atol=0.0, restart=restart, maxiter=maxiter, callback=cb,
                                              callback_type='pr_norm')
        self.iterations = count[0]
        if info != 0:
            raise Exception('NeumannCorrection: GMRES on the boundary system did not converge (info = %d)' % info)
        return tau

    def boundary_normal_derivative_of_sources(self, tau):
        """d/dn of the correction on every boundary (the quantity GMRES matched to the data)"""
        return self._apply(np.asarray(tau, dtype=float).ravel())

    def __call__(self, ue, bc_n, **kwargs):
        bn = self.solver.get_boundary_normal_derivatives(ue)
        g = np.asarray(bc_n, dtype=float).ravel() - np.asarray(bn, dtype=float).ravel()
        self.tau = self.densities(g, **kwargs)
        return self._add_field(ue, self.tau)


class StokesDirichletCorrection(object):
    """
    Homogeneous (no-slip / prescribed velocity) correction of the Stokes solve, examples/interior_stokes_with_multi.py:
    98-118: after `u, v, p = solver(fu, fv)` add the Stokes flow whose boundary velocity is bc - (u, v)|_bdy.

    The reference solves (D - I/2 + N) tau = g with pybie2d's singular double-layer form and a pressure fix, then
    tau -> Stokes_QFS sources -> stokeslet sums on `ebdyc.grid_and_radial_pts`.  Here, as in DirichletCorrection, the
    flow is represented by stokeslets on the QFS source curve of every boundary, obtained from the boundary velocity
    by the QFS boundary -> source operator:

        tau_i (both components on boundary i),  sigma_i = u2s_i(tau_i)
        tau_i + sum_{j != i} S(bdy_i <- src_j) sigma_j = bc_i - (u, v)|_i

    one boundary: tau = g (the reference's example); several: GMRES on the block system.  Sums: k_layer_stokes through
    the solver's own Layer_Apply.  The pressure is defined up to a constant -- normalise as the example does (:120-125).
    """
    def __init__(self, solver, tol=1e-13):
        self.solver = solver
        e = self.ebdyc = solver.ebdyc
        self.qfs = [Stokes_QFS(ebdy.bdy, ebdy.interior, True, True, tol=tol) for ebdy in e]
        self.nb = [ebdy.bdy.N for ebdy in e]
        self.off = np.concatenate([[0], np.cumsum([2*n for n in self.nb])])
        self.iterations = 0

    def _sources(self, tau):
        return [q.u2s(tau[self.off[i]:self.off[i + 1]]).reshape(2, -1) for i, q in enumerate(self.qfs)]

    def _apply(self, tau):
        sig = self._sources(tau)
        out = np.array(tau, dtype=float)
        for i, ebdy in enumerate(self.ebdyc):
            for j, q in enumerate(self.qfs):
                if i != j:
                    cu, cv, _ = self.solver.Layer_Apply(q.source, ebdy.bdy, sig[j])
                    out[self.off[i]:self.off[i + 1]] += np.concatenate([cu, cv])
        return out

    def densities(self, g, tol=1e-13, restart=60, maxiter=10):
        if len(self.qfs) == 1:
            self.iterations = 0
            return g
        count = [0]
        def cb(_):
            count[0] += 1
        A = scipy.sparse.linalg.LinearOperator((g.size, g.size), matvec=self._apply, dtype=float)
        tau, info = scipy.sparse.linalg.gmres(A, g, rtol=tol, atol=0.0, restart=restart, maxiter=maxiter, callback=cb,
                                              callback_type='pr_norm')
        self.iterations = count[0]
        if info != 0:
            raise Exception('StokesDirichletCorrection: GMRES on the boundary system did not converge (info = %d)' % info)
        return tau

    def __call__(self, u, v, p, bcu, bcv, **kwargs):
        """bcu, bcv: prescribed boundary velocity, concatenated over the boundaries -> corrected (u, v, p)"""
        bu = np.asarray(self.solver.get_boundary_values(u), dtype=float)
        bv = np.asarray(self.solver.get_boundary_values(v), dtype=float)
        gu, gv = np.asarray(bcu, dtype=float).ravel() - bu, np.asarray(bcv, dtype=float).ravel() - bv
        boff = np.concatenate([[0], np.cumsum(self.nb)])
        g = np.concatenate([np.concatenate([gu[boff[i]:boff[i + 1]], gv[boff[i]:boff[i + 1]]]) for i in range(len(self.nb))])
        tau = self.densities(g, **kwargs)
        pts = self.ebdyc.grid_and_radial_pts
        out = [np.array(np.asarray(a), dtype=float) for a in (u, v, p)]
        for q, sig in zip(self.qfs, self._sources(tau)):
            for a, c in zip(out, self.solver.Layer_Apply(q.source, pts, sig)):
                a += c
        return tuple(EmbeddedFunction(self.ebdyc, linear_data=a) for a in out)
