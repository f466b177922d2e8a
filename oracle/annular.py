"""
ORACLE (test infrastructure, CPU, NumPy) -- never imported by the product.

Restatement of the reference's annular spectral solvers for the scalar
problems:

  ChebyshevOperators          ipde/annular/annular_full.py:7-50
  ApproximateAnnularGeometry  ipde/annular/annular_full.py:52-85  (ns = n)
                              ipde/annular/annular.py:52-85       (ns = n-1, Nyquist dropped)
  RealAnnularGeometry         ipde/annular/annular_full.py:87-108
  fscalar_laplacian           ipde/annular/modified_helmholtz.py:25-38
  AnnularModifiedHelmholtzSolver  ipde/annular/modified_helmholtz.py:90-203
  AnnularPoissonSolver        ipde/annular/poisson.py:3-21
  right_gmres                 personal_utilities.scipy_gmres (third party, absent,
                              unpinned): right-preconditioned restarted GMRES with a
                              relative-residual stopping test; restated here as
                              CGS2 Arnoldi + Givens rotations.

Pinned: tests/golden/annular_mh.npz holds inputs and outputs of the reference's
own AnnularModifiedHelmholtzSolver / AnnularPoissonSolver (generated in the build
container by oracle/make_golden.py importing /root/reference); tests/test_oracle.py
checks this file against them.
"""
import numpy as np

cheb = np.polynomial.chebyshev


def chebyshev_nodes(lb, ub, order):
    xc, _ = cheb.chebgauss(order)
    xc = xc[::-1]
    rat = (ub - lb)/2.0
    return xc, (xc + 1.0)*rat + lb, rat


class ChebyshevOperators(object):
    def __init__(self, M, rat):
        self.M = M
        xc0, _ = cheb.chebgauss(M)
        xc1, _ = cheb.chebgauss(M - 1)
        xc2, _ = cheb.chebgauss(M - 2)
        self.V0 = cheb.chebvander(xc0, M - 1)
        self.V1 = cheb.chebvander(xc1, M - 2)
        self.V2 = cheb.chebvander(xc2, M - 3)
        self.VI0 = np.linalg.inv(self.V0)
        self.VI1 = np.linalg.inv(self.V1)
        self.VI2 = np.linalg.inv(self.V2)
        DC01 = cheb.chebder(np.eye(M))/rat
        DC12 = cheb.chebder(np.eye(M - 1))/rat
        DC00 = np.vstack([DC01, np.zeros(M)])
        self.D00 = self.V0.dot(DC00.dot(self.VI0))
        self.D01 = self.V1.dot(DC01.dot(self.VI0))
        self.D12 = self.V2.dot(DC12.dot(self.VI1))
        self.ibc_dirichlet = cheb.chebvander(1, M - 1).dot(self.VI0)
        self.obc_dirichlet = cheb.chebvander(-1, M - 1).dot(self.VI0)
        self.ibc_neumann = self.ibc_dirichlet.dot(self.D00)
        self.obc_neumann = self.obc_dirichlet.dot(self.D00)
        t = np.zeros([M - 1, M]); np.fill_diagonal(t, 1.0)
        self.R01 = self.V1.dot(t.dot(self.VI0))
        t = np.zeros([M - 2, M - 1]); np.fill_diagonal(t, 1.0)
        self.R12 = self.V2.dot(t.dot(self.VI1))
        self.R02 = self.R12.dot(self.R01)
        t = np.zeros([M, M - 1]); np.fill_diagonal(t, 1.0)
        self.P10 = self.V0.dot(t.dot(self.VI1))


class ApproximateAnnularGeometry(object):
    def __init__(self, n, M, width, approx_r, drop_nyquist=False):
        self.n, self.M = n, M
        self.radius, self.width = approx_r, width
        self.n2 = int(n/2)
        self.k = np.fft.fftfreq(n, 1.0/n)
        if drop_nyquist:
            self.ns = n - 1
            self.ks = np.concatenate([self.k[:self.n2], self.k[self.n2 + 1:]])
        else:
            self.ns = n
            self.ks = self.k
        self.iks = 1j*self.ks
        _, self.rv0, rat0 = chebyshev_nodes(-width, 0.0, M)
        _, self.rv1, _ = chebyshev_nodes(-width, 0.0, M - 1)
        _, self.rv2, _ = chebyshev_nodes(-width, 0.0, M - 2)
        self.ratio = -rat0
        self.approx_psi0 = approx_r + self.rv0
        self.approx_psi1 = approx_r + self.rv1
        self.approx_psi2 = approx_r + self.rv2
        self.approx_inv_psi0 = 1.0/self.approx_psi0
        self.approx_inv_psi1 = 1.0/self.approx_psi1
        self.approx_inv_psi2 = 1.0/self.approx_psi2
        self.CO = ChebyshevOperators(M, self.ratio)


class RealAnnularGeometry(object):
    def __init__(self, speed, curvature, AAG):
        n = curvature.shape[0]
        k = np.fft.fftfreq(n, 1.0/n)
        dt_curvature = np.fft.ifft(np.fft.fft(curvature)*1j*k).real
        rv0, rv1, rv2 = AAG.rv0, AAG.rv1, AAG.rv2
        self.psi0 = speed*(1 + rv0[:, None]*curvature)
        self.psi1 = speed*(1 + rv1[:, None]*curvature)
        self.psi2 = speed*(1 + rv2[:, None]*curvature)
        self.inv_psi0 = 1.0/self.psi0
        self.inv_psi1 = 1.0/self.psi1
        self.inv_psi2 = 1.0/self.psi2
        self.DR_psi2 = speed*curvature*np.ones(rv2[:, None].shape)
        idenom2 = 1.0/(speed*(1 + rv2[:, None]*curvature)**3)
        self.ipsi_DR_ipsi_DT_psi2 = dt_curvature*idenom2
        self.ipsi_DT_ipsi_DR_psi2 = dt_curvature*idenom2


def right_gmres(apply_A, b, apply_M, tol=1e-5, restart=20, maxiter=None, verbose=False):
    """
    Solve A x = b with right preconditioning (A M y = b, x = M y), restarted GMRES.
    Stops when ||b - A x|| <= tol*||b||.  `maxiter` counts restart cycles (SciPy's
    convention).  Returns (x, info, list of relative residuals, one per inner iteration).
    """
    n = b.shape[0]
    if restart is None:
        restart = 20
    restart = min(restart, n)
    if maxiter is None:
        maxiter = 10*n
    bnorm = np.linalg.norm(b)
    resids = []
    x = np.zeros(n, dtype=complex)
    if bnorm == 0.0:
        return x, 0, resids
    y = np.zeros(n, dtype=complex)     # solution in the preconditioned variable
    for cycle in range(maxiter):
        r = b - apply_A(apply_M(y)) if cycle > 0 else b.astype(complex)
        beta = np.linalg.norm(r)
        if beta <= tol*bnorm:
            return apply_M(y), 0, resids
        V = np.zeros((restart + 1, n), dtype=complex)
        H = np.zeros((restart + 1, restart), dtype=complex)
        cs = np.zeros(restart, dtype=complex)
        sn = np.zeros(restart, dtype=complex)
        g = np.zeros(restart + 1, dtype=complex)
        g[0] = beta
        V[0] = r/beta
        kdone = 0
        converged = False
        for j in range(restart):
            w = apply_A(apply_M(V[j]))
            # classical Gram-Schmidt, twice
            h = np.conj(V[:j + 1]).dot(w)
            w = w - h.dot(V[:j + 1])
            h2 = np.conj(V[:j + 1]).dot(w)
            w = w - h2.dot(V[:j + 1])
            h = h + h2
            hn = np.linalg.norm(w)
            H[:j + 1, j] = h
            H[j + 1, j] = hn
            if hn > 0:
                V[j + 1] = w/hn
            for i in range(j):
                t = cs[i]*H[i, j] + sn[i]*H[i + 1, j]
                H[i + 1, j] = -np.conj(sn[i])*H[i, j] + cs[i]*H[i + 1, j]
                H[i, j] = t
            a, bb = H[j, j], H[j + 1, j]
            den = np.sqrt(abs(a)**2 + abs(bb)**2)
            cs[j] = abs(a)/den if a != 0 else 0.0
            sn[j] = (a/abs(a))*np.conj(bb)/den if a != 0 else 1.0
            H[j, j] = cs[j]*a + sn[j]*bb
            H[j + 1, j] = 0.0
            g[j + 1] = -np.conj(sn[j])*g[j]
            g[j] = cs[j]*g[j]
            kdone = j + 1
            rel = abs(g[j + 1])/bnorm
            resids.append(rel)
            if verbose:
                print('   gmres it', len(resids), 'rel res', rel)
            if rel <= tol or hn == 0.0:
                converged = True
                break
        yk = np.linalg.solve(np.triu(H[:kdone, :kdone]), g[:kdone])
        y = y + yk.dot(V[:kdone])
        if converged:
            return apply_M(y), 0, resids
    return apply_M(y), 1, resids


class AnnularModifiedHelmholtzSolver(object):
    """
    (k^2 - Laplacian) u = f on the annulus, Robin conditions
    ia*u + ib*u_r = ig at the inner radius, oa*u + ob*u_r = og at the outer one.
    """
    def __init__(self, AAG, k, ia=1.0, ib=0.0, oa=1.0, ob=0.0):
        self.AAG = AAG
        self.ia, self.ib, self.oa, self.ob = ia, ib, oa, ob
        self.k = k
        self.M, self.ns, self.n = AAG.M, AAG.ns, AAG.n
        self.small_shape = (self.M, self.ns)
        self.shape = (self.M, self.n)
        self._construct()

    def _construct(self):
        AAG, CO = self.AAG, self.AAG.CO
        M, ns = self.M, self.ns
        # radial part of the concentric-circle operator (independent of the mode)
        A0 = AAG.approx_inv_psi2[:, None]*CO.D12.dot(AAG.approx_psi1[:, None]*CO.D01)
        A1 = CO.R12.dot(AAG.approx_inv_psi1[:, None]*CO.R01)
        K = np.empty((ns, M, M), dtype=float)
        K[:, :M - 2] = self.k**2*CO.R02 - (A0[None] - (AAG.ks**2)[:, None, None]*A1[None])
        K[:, M - 2] = self.ia*CO.ibc_dirichlet + self.ib*CO.ibc_neumann
        K[:, M - 1] = self.oa*CO.obc_dirichlet + self.ob*CO.obc_neumann
        self.Stacked_KINVS = np.linalg.inv(K)

    def _preconditioner(self, fh):
        # per-mode dense mat-vec on the mode-minor vector (modified_helmholtz.py:68-88)
        x = fh.reshape(self.small_shape)
        return np.einsum('ijk,ki->ji', self.Stacked_KINVS, x).ravel()

    def _fm(self, fh, m):
        return np.fft.fft(m*np.fft.ifft(fh))

    def _laplacian(self, uh):
        CO, AAG, RAG = self.AAG.CO, self.AAG, self.RAG
        iks = AAG.iks
        uh_t = CO.R01.dot(uh*iks)
        uh_tt = CO.R12.dot(self._fm(uh_t, RAG.inv_psi1)*iks)
        uh_rr = CO.D12.dot(self._fm(CO.D01.dot(uh), RAG.psi1))
        return self._fm(uh_rr + uh_tt, RAG.inv_psi2)

    def _apply(self, uh):
        CO = self.AAG.CO
        uh = uh.reshape(self.small_shape)
        fuh = self.k**2*CO.R02.dot(uh) - self._laplacian(uh)
        ibc = (self.ia*CO.ibc_dirichlet + self.ib*CO.ibc_neumann).dot(uh)
        obc = (self.oa*CO.obc_dirichlet + self.ob*CO.obc_neumann).dot(uh)
        return np.concatenate([fuh.ravel(), ibc.ravel(), obc.ravel()])

    def solve(self, RAG, f, ig, og, ia=None, ib=None, oa=None, ob=None, verbose=False, **kwargs):
        self.RAG = RAG
        self.ia = ia if ia is not None else self.ia
        self.ib = ib if ib is not None else self.ib
        self.oa = oa if oa is not None else self.oa
        self.ob = ob if ob is not None else self.ob
        R02 = self.AAG.CO.R02
        ff = np.concatenate([R02.dot(f).ravel(), np.asarray(ig).ravel(), np.asarray(og).ravel()])
        ffh = np.fft.fft(ff.reshape(self.shape)).ravel()
        x, info, resids = right_gmres(self._apply, ffh, self._preconditioner, verbose=verbose, **kwargs)
        self.iterations_last_call = len(resids)
        self.residuals_last_call = resids
        return np.fft.ifft(x.reshape(self.small_shape)).real


class AnnularPoissonSolver(AnnularModifiedHelmholtzSolver):
    """Laplacian u = f (annular/poisson.py:16-21: k = 0 and f -> -f)."""
    def __init__(self, AAG, ia=1.0, ib=0.0, oa=1.0, ob=0.0):
        super().__init__(AAG, 0.0, ia=ia, ib=ib, oa=oa, ob=ob)

    def solve(self, RAG, f, ig, og, **kwargs):
        return super().solve(RAG, -f, ig, og, **kwargs)
