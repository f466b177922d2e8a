"""
Annular geometry/operators (host setup) and the annular solver classes (device solve).

API mirror of the reference's ipde/annular package:
  ChebyshevOperators            annular_full.py:7-50
  ApproximateAnnularGeometry    annular_full.py:52-85 (all n modes) / annular.py:52-85 (Nyquist dropped)
  RealAnnularGeometry           annular_full.py:87-108
  AnnularModifiedHelmholtzSolver(AAG, k, ia, ib, oa, ob).solve(RAG, f, ig, og, **kw)
                                modified_helmholtz.py:90-203
  AnnularPoissonSolver          poisson.py:3-21
The small dense matrices and the per-mode preconditioner inverses are built here in NumPy
(setup, done once per solver) and uploaded; `solve` runs entirely on the device: theta-FFTs,
operator apply, batched per-mode preconditioner and GMRES (csrc/annular.cu, csrc/gmres.cu).
"""
import ctypes
import numpy as np
import torch
from ._lib import lib
from .device import to_dev, empty, ptr, hptr, stream, require_cuda

cheb = np.polynomial.chebyshev


def _nodes(lb, ub, order):
    xc, _ = cheb.chebgauss(order)
    xc = xc[::-1]
    rat = (ub - lb)/2.0
    return (xc + 1.0)*rat + lb, rat


class ChebyshevOperators(object):
    """Differentiation / restriction / boundary rows between Chebyshev-Gauss grids of M, M-1, M-2 points."""
    def __init__(self, M, rat):
        self.M = M
        x0, _ = cheb.chebgauss(M)
        x1, _ = cheb.chebgauss(M - 1)
        x2, _ = cheb.chebgauss(M - 2)
        self.V0, self.V1, self.V2 = cheb.chebvander(x0, M - 1), cheb.chebvander(x1, M - 2), cheb.chebvander(x2, M - 3)
        self.VI0, self.VI1, self.VI2 = np.linalg.inv(self.V0), np.linalg.inv(self.V1), np.linalg.inv(self.V2)
        DC01 = cheb.chebder(np.eye(M))/rat
        DC12 = cheb.chebder(np.eye(M - 1))/rat
        DC00 = np.vstack([DC01, np.zeros(M)])
        self.D00 = self.V0 @ DC00 @ self.VI0
        self.D01 = self.V1 @ DC01 @ self.VI0
        self.D12 = self.V2 @ DC12 @ self.VI1
        self.ibc_dirichlet = cheb.chebvander(1, M - 1) @ self.VI0
        self.obc_dirichlet = cheb.chebvander(-1, M - 1) @ self.VI0
        self.ibc_neumann = self.ibc_dirichlet @ self.D00
        self.obc_neumann = self.obc_dirichlet @ self.D00
        self.R01 = self.V1 @ np.eye(M - 1, M) @ self.VI0
        self.R12 = self.V2 @ np.eye(M - 2, M - 1) @ self.VI1
        self.R02 = self.R12 @ self.R01
        self.P10 = self.V0 @ np.eye(M, M - 1) @ self.VI1


class ApproximateAnnularGeometry(object):
    """
    Concentric-circle stand-in for the annulus (used by the preconditioner).
    drop_nyquist=False is the reference's annular_full.py (scalar solvers), True its annular.py (Stokes).
    """
    def __init__(self, n, M, width, approx_r, drop_nyquist=False):
        self.n, self.M = n, M
        self.radius, self.width = approx_r, width
        self.radial_h = width/M
        self.tangent_h = 2*np.pi/n
        self.n2 = int(n/2)
        self.k = np.fft.fftfreq(n, 1.0/n)
        if drop_nyquist:
            self.ns = n - 1
            self.ks = np.concatenate([self.k[:self.n2], self.k[self.n2 + 1:]])
        else:
            self.ns = n
            self.ks = self.k
        self.iks = 1j*self.ks
        self.rv0, rat0 = _nodes(-width, 0.0, M)
        self.rv1, _ = _nodes(-width, 0.0, M - 1)
        self.rv2, _ = _nodes(-width, 0.0, M - 2)
        self.ratio = -rat0
        self.approx_psi0 = approx_r + self.rv0
        self.approx_psi1 = approx_r + self.rv1
        self.approx_psi2 = approx_r + self.rv2
        self.approx_inv_psi0 = 1.0/self.approx_psi0
        self.approx_inv_psi1 = 1.0/self.approx_psi1
        self.approx_inv_psi2 = 1.0/self.approx_psi2
        self.CO = ChebyshevOperators(M, self.ratio)


class RealAnnularGeometry(object):
    """psi = speed*(1 + r*curvature) on the three radial grids, and the metric terms Stokes needs."""
    def __init__(self, speed, curvature, AAG):
        n = curvature.shape[0]
        k = np.fft.fftfreq(n, 1.0/n)
        dt_curvature = np.fft.ifft(np.fft.fft(curvature)*1j*k).real
        self.psi0 = speed*(1 + AAG.rv0[:, None]*curvature)
        self.psi1 = speed*(1 + AAG.rv1[:, None]*curvature)
        self.psi2 = speed*(1 + AAG.rv2[:, None]*curvature)
        self.inv_psi0, self.inv_psi1, self.inv_psi2 = 1.0/self.psi0, 1.0/self.psi1, 1.0/self.psi2
        self.DR_psi2 = speed*curvature*np.ones(AAG.rv2[:, None].shape)
        idenom2 = 1.0/(speed*(1 + AAG.rv2[:, None]*curvature)**3)
        self.ipsi_DR_ipsi_DT_psi2 = dt_curvature*idenom2
        self.ipsi_DT_ipsi_DR_psi2 = dt_curvature*idenom2


class AnnularModifiedHelmholtzSolver(object):
    """
    Solves (k^2 - Laplacian) u = f on the annulus with
        ia*u + ib*u_r = ig at the inner radius,  oa*u + ob*u_r = og at the outer radius.
    `solve` accepts host (numpy) or device (torch) arrays and returns the same kind.
    """
    def __init__(self, AAG, k, ia=1.0, ib=0.0, oa=1.0, ob=0.0):
        require_cuda()
        self.AAG = AAG
        self.ia, self.ib, self.oa, self.ob = ia, ib, oa, ob
        self.k = k
        self.M, self.ns, self.n = AAG.M, AAG.ns, AAG.n
        if self.ns != self.n:
            raise Exception('scalar annular solvers use the full-mode geometry (annular_full)')
        self.small_shape = (self.M, self.ns)
        self.shape = (self.M, self.n)
        self._construct()
        self._plans = {}
        self.iterations_last_call = None

    def _construct(self):
        AAG, CO, M = self.AAG, self.AAG.CO, self.M
        A0 = AAG.approx_inv_psi2[:, None]*(CO.D12 @ (AAG.approx_psi1[:, None]*CO.D01))
        A1 = CO.R12 @ (AAG.approx_inv_psi1[:, None]*CO.R01)
        K = np.empty((self.ns, M, M))
        K[:, :M - 2] = self.k**2*CO.R02 - (A0[None] - (AAG.ks**2)[:, None, None]*A1[None])
        K[:, M - 2] = self.ia*CO.ibc_dirichlet + self.ib*CO.ibc_neumann
        K[:, M - 1] = self.oa*CO.obc_dirichlet + self.ob*CO.obc_neumann
        self.Stacked_KINVS = np.linalg.inv(K)

    def _get_plan(self, RAG):
        key = (id(RAG), self.ia, self.ib, self.oa, self.ob)
        # one device plan per (geometry, boundary rows): an annular solver shared between boundaries of the same size
        # (helpers= / AS_list= reuse, as in the reference) keeps a plan for each instead of rebuilding it every solve
        hit = self._plans.get(key)
        if hit is not None:
            return hit[0]
        if len(self._plans) >= 16:
            for h_old, _ in self._plans.values():
                lib().ipde_annular_plan_destroy(h_old)
            self._plans.clear()
        CO, AAG = self.AAG.CO, self.AAG
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        ibc = c((self.ia*CO.ibc_dirichlet + self.ib*CO.ibc_neumann)[0])
        obc = c((self.oa*CO.obc_dirichlet + self.ob*CO.obc_neumann)[0])
        kinv = c(np.transpose(self.Stacked_KINVS, (1, 2, 0)))     # (M, M, n): mode-minor
        arrs = [c(CO.D01), c(CO.D12), c(CO.R01), c(CO.R12), c(CO.R02), ibc, obc, c(AAG.ks),
                c(RAG.psi1), c(RAG.inv_psi1), c(RAG.inv_psi2), kinv]
        h = ctypes.c_void_p()
        lib().ipde_annular_plan_create(self.n, self.M, float(self.k), *[hptr(a) for a in arrs], ctypes.byref(h))
        self._plans[key] = (h, RAG)          # the RAG reference keeps id(RAG) unique while the plan lives
        return h

    def __del__(self):
        try:
            for h_old, _ in self._plans.values():
                lib().ipde_annular_plan_destroy(h_old)
        except Exception:
            pass

    _f_sign = 1.0

    def solve(self, RAG, f, ig, og, ia=None, ib=None, oa=None, ob=None, verbose=False,
              tol=1e-5, restart=20, maxiter=None, **kwargs):
        self.RAG = RAG
        self.ia = ia if ia is not None else self.ia
        self.ib = ib if ib is not None else self.ib
        self.oa = oa if oa is not None else self.oa
        self.ob = ob if ob is not None else self.ob
        plan = self._get_plan(RAG)
        on_host = not torch.is_tensor(f)
        fd = to_dev(np.asarray(f, dtype=float)) if on_host else f
        igd = None if ig is None else (to_dev(np.asarray(ig, dtype=float)) if not torch.is_tensor(ig) else ig)
        ogd = None if og is None else (to_dev(np.asarray(og, dtype=float)) if not torch.is_tensor(og) else og)
        out = empty(self.shape)
        iters = ctypes.c_int(0)
        relres = ctypes.c_double(0.0)
        if restart is None:
            restart = 20
        if maxiter is None:
            maxiter = 10*self.M*self.n
        lib().ipde_annular_solve(plan, ptr(fd), ptr(igd), ptr(ogd), self._f_sign, float(tol), int(restart),
                                 int(maxiter), ptr(out), ctypes.byref(iters), ctypes.byref(relres), stream())
        self.iterations_last_call = iters.value
        self.relative_residual_last_call = relres.value
        if verbose:
            print('GMRES took:', iters.value, 'iterations.')
        return out.cpu().numpy() if on_host else out

    # exposed for tests: operator and preconditioner applied to a complex (M*n) host vector
    def _apply(self, uh):
        plan = self._get_plan(self.RAG)
        x = to_dev(np.asarray(uh, dtype=complex))
        y = torch.empty_like(x)
        lib().ipde_annular_apply(plan, ptr(x), ptr(y), stream())
        return y.cpu().numpy()

    def _optim_preconditioner(self, fh):
        plan = self._get_plan(self.RAG)
        x = to_dev(np.asarray(fh, dtype=complex))
        y = torch.empty_like(x)
        lib().ipde_annular_precond(plan, ptr(x), ptr(y), stream())
        return y.cpu().numpy()


class AnnularPoissonSolver(AnnularModifiedHelmholtzSolver):
    """Laplacian u = f  (reference annular/poisson.py: k = 0 and f -> -f)."""
    _f_sign = -1.0

    def __init__(self, AAG, ia=1.0, ib=0.0, oa=1.0, ob=0.0):
        super().__init__(AAG, 0.0, ia=ia, ib=ib, oa=oa, ob=ob)


class AnnularStokesSolver(object):
    """
    -mu Lap u + grad p = f, div u = 0 on the annulus, in (r, t) components, Dirichlet data at both radii
    (reference ipde/annular/stokes.py:73-541).  Works on the Nyquist-dropped geometry
    (ApproximateAnnularGeometry(..., drop_nyquist=True), the reference's annular.py).
    `solve` accepts host or device arrays and returns the same kind.
    """
    def __init__(self, AAG, mu=1.0):
        require_cuda()
        self.AAG, self.mu = AAG, mu
        self.M, self.ns, self.n = AAG.M, AAG.ns, AAG.n
        if self.ns != self.n - 1:
            raise Exception('the annular Stokes solver works on the Nyquist-dropped geometry')
        self.u_shape = (self.M, self.n)
        self._construct()
        self._plans = {}
        self.iterations_last_call = None

    def _construct(self):
        """per-mode (3M-1)^2 complex collocation matrices of the concentric-circle operator, inverted (:127-153)"""
        AAG, CO, M, ns, mu = self.AAG, self.AAG.CO, self.M, self.ns, self.mu
        ap0, ap1 = AAG.approx_psi0, AAG.approx_psi1
        ai1, ai2 = AAG.approx_inv_psi1, AAG.approx_inv_psi2
        ks = AAG.ks
        D01, D12, R01, R12, R02 = CO.D01, CO.D12, CO.R01, CO.R12, CO.R02
        A0 = ai2[:, None]*(D12 @ (ap1[:, None]*D01))
        A1 = R12 @ (ai1[:, None]*R01)
        K = np.zeros((ns, 3*M - 1, 3*M - 1), dtype=complex)
        LL = A0[None] - (ks**2)[:, None, None]*A1[None]
        diag = -LL + (ai2**2)[:, None]*R02
        cross = (2*ai2**2)[:, None]*R02
        ik = 1j*ks[:, None, None]
        K[:, 0:M - 2, 0:M] = mu*diag
        K[:, 0:M - 2, M:2*M] = mu*ik*cross[None]
        K[:, 0:M - 2, 2*M:] = D12
        K[:, M - 2, 0:M] = CO.ibc_dirichlet[0]
        K[:, M - 1, 0:M] = CO.obc_dirichlet[0]
        K[:, M:2*M - 2, 0:M] = -mu*ik*cross[None]
        K[:, M:2*M - 2, M:2*M] = mu*diag
        K[:, M:2*M - 2, 2*M:] = ik*(ai2[:, None]*R12)[None]
        K[:, 2*M - 2, M:2*M] = CO.ibc_dirichlet[0]
        K[:, 2*M - 1, M:2*M] = CO.obc_dirichlet[0]
        K[:, 2*M:, 0:M] = ai1[:, None]*(D01*ap0[None, :])
        K[:, 2*M:, M:2*M] = ik*(ai1[:, None]*R01)[None]
        K[0, 2*M:, 2*M:] += CO.VI1[0]
        self.Stacked_KINVS = np.linalg.inv(K)

    def _get_plan(self, RAG):
        key = id(RAG)
        # one device plan per (geometry, boundary rows): an annular solver shared between boundaries of the same size
        # (helpers= / AS_list= reuse, as in the reference) keeps a plan for each instead of rebuilding it every solve
        hit = self._plans.get(key)
        if hit is not None:
            return hit[0]
        if len(self._plans) >= 16:
            for h_old, _ in self._plans.values():
                lib().ipde_annular_stokes_plan_destroy(h_old)
            self._plans.clear()
        CO, AAG, M, n = self.AAG.CO, self.AAG, self.M, self.n
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        ks = AAG.k.copy()
        ks[n//2] = 0.0
        R = 3*M - 1
        kinv = np.zeros((R, R, n), dtype=complex)
        cols = np.concatenate([np.arange(n//2), np.arange(n//2 + 1, n)])
        kinv[:, :, cols] = np.transpose(self.Stacked_KINVS, (1, 2, 0))
        kinv = np.ascontiguousarray(kinv)
        combo1 = 2*RAG.DR_psi2*RAG.inv_psi2**2
        combo2 = RAG.DR_psi2**2*RAG.inv_psi2**2
        arrs = [c(CO.D01), c(CO.D12), c(CO.R01), c(CO.R12), c(CO.R02), c(CO.P10), c(CO.ibc_dirichlet[0]),
                c(CO.obc_dirichlet[0]), c(CO.VI1[0]), c(ks), c(RAG.psi0), c(RAG.psi1), c(RAG.inv_psi1), c(RAG.inv_psi2),
                c(combo1), c(combo2), c(RAG.ipsi_DR_ipsi_DT_psi2)]
        h = ctypes.c_void_p()
        lib().ipde_annular_stokes_plan_create(n, M, float(self.mu), *[hptr(a) for a in arrs], hptr(kinv), ctypes.byref(h))
        self._plans[key] = (h, RAG)          # the RAG reference keeps id(RAG) unique while the plan lives
        return h

    def __del__(self):
        try:
            for h_old, _ in self._plans.values():
                lib().ipde_annular_stokes_plan_destroy(h_old)
        except Exception:
            pass

    def solve(self, RAG, fr, ft, irg, itg, org, otg, verbose=False, tol=1e-5, restart=20, maxiter=None, **kwargs):
        self.RAG = RAG
        plan = self._get_plan(RAG)
        on_host = not torch.is_tensor(fr)
        dv = lambda a: None if a is None else (a if torch.is_tensor(a) else to_dev(np.asarray(a, dtype=float)))
        frd, ftd = dv(fr), dv(ft)
        bcs = [dv(b) for b in (irg, itg, org, otg)]
        ur, ut, pr = empty(self.u_shape), empty(self.u_shape), empty(self.u_shape)
        iters = ctypes.c_int(0)
        relres = ctypes.c_double(0.0)
        if restart is None:
            restart = 20
        if maxiter is None:
            maxiter = 100
        lib().ipde_annular_stokes_solve(plan, ptr(frd), ptr(ftd), ptr(bcs[0]), ptr(bcs[1]), ptr(bcs[2]), ptr(bcs[3]),
                                        float(tol), int(restart), int(maxiter), ptr(ur), ptr(ut), ptr(pr),
                                        ctypes.byref(iters), ctypes.byref(relres), stream())
        self.iterations_last_call = iters.value
        self.relative_residual_last_call = relres.value
        if verbose:
            print('GMRES took:', iters.value, 'iterations.')
        if on_host:
            return ur.cpu().numpy(), ut.cpu().numpy(), pr.cpu().numpy()
        return ur, ut, pr

    # ---- exposed for tests: vectors in the reference's layout (ns = n-1 columns per row block, raveled)
    def _to_dev_layout(self, v):
        M, n, ns = self.M, self.n, self.ns
        R = 3*M - 1
        full = np.zeros((R, n), dtype=complex)
        cols = np.concatenate([np.arange(n//2), np.arange(n//2 + 1, n)])
        full[:, cols] = np.asarray(v, dtype=complex).reshape(R, ns)
        return to_dev(full)

    def _from_dev_layout(self, y):
        n = self.n
        cols = np.concatenate([np.arange(n//2), np.arange(n//2 + 1, n)])
        return y.cpu().numpy()[:, cols].ravel()

    def _apply_optim_real(self, uuh):
        plan = self._get_plan(self.RAG)
        x = self._to_dev_layout(uuh)
        y = torch.empty_like(x)
        lib().ipde_annular_stokes_apply(plan, ptr(x), ptr(y), stream())
        return self._from_dev_layout(y)

    def _preconditioner(self, ffh):
        plan = self._get_plan(self.RAG)
        x = self._to_dev_layout(ffh)
        y = torch.empty_like(x)
        lib().ipde_annular_stokes_precond(plan, ptr(x), ptr(y), stream())
        return self._from_dev_layout(y)
