"""
Quadrature by fundamental solutions (QFS) -- setup on the host, dense NumPy/SciPy.

The reference takes `Laplace_QFS`, `Modified_Helmholtz_QFS`, `Stokes_QFS` and
`QFS_Boundary` from the un-vendored, unpinned `qfs` package (call sites
ipde/solvers/internals/poisson.py:23-26, modified_helmholtz.py:24-27, stokes.py:21-24,
used at internals/scalar.py:87-88,101-113 and internals/vector.py:124-125,150-160).
This is a restatement of the published algorithm (Stein & Barnett, "Quadrature by
fundamental solutions", Adv. Comput. Math. 2022), in its desingularised form:

  * sources: M = rho*N point singularities on the curve offset a distance
    alpha_s*h from the boundary, on the side *away* from where the potential is wanted;
  * check points: M points on the curve offset alpha_c*h *into* the evaluation side;
  * the layer potential S[tau_s] + D[tau_d] is evaluated on the check curve by the
    periodic trapezoid rule on a rho_f-times upsampled boundary (spectrally accurate at
    that distance), and the source strengths solve  E sigma = C tau  with E the
    check<-source matrix.  E is exponentially ill-conditioned, so a pre-multiplied
    E^-1 C loses all accuracy; the solve is kept factored, E = U S V^T (truncated SVD):
    sigma = V (S^+ ((U^T C) tau)) -- two well-conditioned mat-vecs and a diagonal scaling,
    which is backward stable, gives the minimum-norm strengths, and is insensitive to
    summation order (what makes CPU-oracle / GPU parity at 1e-11 possible).
  * `u2s(ub)`: strengths whose potential equals `ub` on the boundary nodes
    (collocation on the boundary itself, upsampled to M points).

What the per-solve path needs from an instance -- `source` (a curve object with x, y,
weights), `__call__([slp, dlp]) -> sigma`, `u2s(ub) -> sigma` -- matches the reference's
usage.  The dense operators (`A = U^T C`, `Sinv`, `V` and their u2s counterparts) are exposed as
arrays so that the CUDA library applies them on the device; the NumPy `__call__` below is the
setup-time / oracle-side statement of the same arithmetic.

Kernel normalisations follow the reference's in-tree pins: Laplace -log r/(2 pi)
(grid_evaluators/laplace_grid_evaluator.py:8-12), modified Helmholtz K0(kr)/(2 pi)
(modified_helmholtz_grid_evaluator.py:8-9), Stokes stokeslet (1/4pi)(-log r I + rr/r^2),
pressure (1/2pi) r.f/r^2 (solvers/internals/stokes_save.py:29-38).
"""
import numpy as np
import scipy.linalg
import scipy.special
from .geometry import GSB, fourier_resample, fourier_resample_matrix

TWO_PI = 2.0*np.pi


# ---------------------------------------------------------------------------
# kernels (dense matrix forms; targets = rows, sources = columns)
def _diffs(tx, ty, sx, sy):
    dx = tx[:, None] - sx[None, :]
    dy = ty[:, None] - sy[None, :]
    return dx, dy


class LaplaceKernel(object):
    name = 'laplace'
    ncomp = 1
    k = 0.0
    def slp(self, tx, ty, sx, sy):
        dx, dy = _diffs(tx, ty, sx, sy)
        return -np.log(dx*dx + dy*dy)/(4.0*np.pi)
    def dlp(self, tx, ty, sx, sy, snx, sny):
        dx, dy = _diffs(tx, ty, sx, sy)
        return (dx*snx[None, :] + dy*sny[None, :])/(dx*dx + dy*dy)/TWO_PI


class ModifiedHelmholtzKernel(object):
    name = 'modified_helmholtz'
    ncomp = 1
    def __init__(self, k):
        self.k = float(k)
    def slp(self, tx, ty, sx, sy):
        dx, dy = _diffs(tx, ty, sx, sy)
        return scipy.special.k0(self.k*np.sqrt(dx*dx + dy*dy))/TWO_PI
    def dlp(self, tx, ty, sx, sy, snx, sny):
        # d/dn_y K0(k|x-y|)/(2 pi) = k K1(k r) (n_y.(x-y))/r /(2 pi)
        dx, dy = _diffs(tx, ty, sx, sy)
        r = np.sqrt(dx*dx + dy*dy)
        return self.k*scipy.special.k1(self.k*r)*(dx*snx[None, :] + dy*sny[None, :])/r/TWO_PI


class StokesKernel(object):
    """2x2 block kernels; vectors are stored [x-components | y-components]."""
    name = 'stokes'
    ncomp = 2
    k = 0.0
    def slp(self, tx, ty, sx, sy):
        dx, dy = _diffs(tx, ty, sx, sy)
        r2 = dx*dx + dy*dy
        L = -0.5*np.log(r2)
        c = 1.0/(4.0*np.pi)
        return np.block([[c*(L + dx*dx/r2), c*(dx*dy/r2)],
                         [c*(dx*dy/r2), c*(L + dy*dy/r2)]])
    def dlp(self, tx, ty, sx, sy, snx, sny):
        # stresslet: u_i = (1/pi) (d_i d_j d_k / r^4) n_k tau_j,  d = x - y
        dx, dy = _diffs(tx, ty, sx, sy)
        r2 = dx*dx + dy*dy
        dn = dx*snx[None, :] + dy*sny[None, :]
        f = dn/(r2*r2)/np.pi
        return np.block([[f*dx*dx, f*dx*dy], [f*dx*dy, f*dy*dy]])


# ---------------------------------------------------------------------------
def offset_curve(c_fine_gsb, dist):
    """points x + dist * n on a GSB (dist > 0: outward for a CCW curve)"""
    return c_fine_gsb.c + dist*c_fine_gsb.normal_c


class QFS_Options(dict):
    pass


def default_parameters(tol):
    """(alpha_s, alpha_c, rho, rho_f) -- distances in units of the boundary spacing h."""
    L = np.log(1.0/max(tol, 1e-15))
    alpha_c = 0.5
    rho = 2.0
    alpha_s = max(1.0, L/(TWO_PI*(rho - 1.0)) - alpha_c)
    rho_f = int(np.ceil((L + 2.0)/(TWO_PI*alpha_c)))
    return alpha_s, alpha_c, rho, max(rho_f, 2)


class QFS(object):
    """
    bdy:       GSB with N nodes
    interior:  True  -> represent the potential on the inside of the curve
    slp, dlp:  which layers the densities passed to __call__ contain (in that order)
    """
    def __init__(self, bdy, interior, slp, dlp, kernel, tol=1e-12,
                 source_upsample_factor=1.0, alpha_s=None, alpha_c=None, rho=None, rho_f=None):
        self.bdy = bdy
        self.interior = interior
        self.has_slp, self.has_dlp = slp, dlp
        self.kernel = kernel
        self.tol = tol
        a_s, a_c, r, r_f = default_parameters(tol)
        self.alpha_s = a_s if alpha_s is None else alpha_s
        self.alpha_c = a_c if alpha_c is None else alpha_c
        self.rho = (r if rho is None else rho)*source_upsample_factor
        self.rho_f = r_f if rho_f is None else rho_f
        self.svd_cut = min(1e-11, max(1e-14, tol))
        self._build()

    def _build(self):
        bdy, ker = self.bdy, self.kernel
        N = bdy.N
        nc = ker.ncomp
        h = bdy.max_h
        M = 2*int(np.ceil(self.rho*N/2.0))
        self.N, self.M = N, M
        # evaluation side: inside -> check curve inward (-n), sources outward (+n)
        sgn = 1.0 if self.interior else -1.0
        # keep the source curve regular: the offset curve x + d n has speed |x'|(1 + kappa d), which must
        # stay well away from zero; pay for a closer source curve with more source points
        worst = np.max(-sgn*bdy.curvature)
        if worst > 0.0 and self.alpha_s*h*worst > 0.4:
            self.alpha_s = 0.4/(worst*h)
            L = np.log(1.0/max(self.tol, 1e-15))
            self.rho = max(self.rho, min(8.0, 1.0 + L/(TWO_PI*(self.alpha_s + self.alpha_c))))
            M = 2*int(np.ceil(self.rho*N/2.0))
            self.M = M
        bM = GSB(c=fourier_resample(bdy.c, M))
        self.source = GSB(c=offset_curve(bM, sgn*self.alpha_s*h))
        self.check = GSB(c=offset_curve(bM, -sgn*self.alpha_c*h))
        src, chk = self.source, self.check
        w_s = np.tile(src.weights, nc)
        UM = fourier_resample_matrix(N, M)
        U_MN = UM if nc == 1 else scipy.linalg.block_diag(*([UM]*nc))
        # check <- boundary densities through the upsampled trapezoid rule
        Nf = self.rho_f*N
        bF = GSB(c=fourier_resample(bdy.c, Nf))
        UF = fourier_resample_matrix(N, Nf)
        w_f = bF.weights
        blocks = []
        if self.has_slp:
            K = ker.slp(chk.x, chk.y, bF.x, bF.y)*np.tile(w_f, nc)[None, :]
            blocks.append(self._times_upsample(K, UF, nc))
        if self.has_dlp:
            K = ker.dlp(chk.x, chk.y, bF.x, bF.y, bF.normal_x, bF.normal_y)*np.tile(w_f, nc)[None, :]
            blocks.append(self._times_upsample(K, UF, nc))
        C = np.hstack(blocks)                                  # (nc*M) x (nlayers*nc*N)
        # check <- source and boundary(M nodes) <- source, each as a truncated SVD  E = U S V^T
        E = ker.slp(chk.x, chk.y, src.x, src.y)*w_s[None, :]
        self.A, self.Sinv, self.V = self._factor(E, C)
        Eb = ker.slp(bM.x, bM.y, src.x, src.y)*w_s[None, :]
        self.Ab, self.Sinv_b, self.Vb = self._factor(Eb, U_MN)

    def _factor(self, E, rhs_op):
        """E = U S V^T  ->  (U^T rhs_op, 1/S with tiny singular values dropped, V), all C-contiguous."""
        U, S, Vt = _svd(E)
        keep = S > self.svd_cut*S[0]
        Sinv = np.where(keep, 1.0/np.where(keep, S, 1.0), 0.0)
        return (np.ascontiguousarray(U.T.dot(rhs_op)), np.ascontiguousarray(Sinv),
                np.ascontiguousarray(Vt.T))

    @staticmethod
    def _times_upsample(K, UF, nc):
        Nf, N = UF.shape
        return np.hstack([K[:, i*Nf:(i + 1)*Nf].dot(UF) for i in range(nc)])

    def __call__(self, tau_list):
        tau = np.concatenate([np.asarray(t, dtype=float).ravel() for t in tau_list])
        return self.V.dot(self.Sinv*self.A.dot(tau))

    def u2s(self, ub):
        return self.Vb.dot(self.Sinv_b*self.Ab.dot(np.asarray(ub, dtype=float).ravel()))


def _svd(E):
    """SVD on the host; large systems go through torch on the GPU when one is present (setup only)."""
    if E.shape[0] > 2500:
        try:
            import torch
            if torch.cuda.is_available():
                U, S, Vh = torch.linalg.svd(torch.from_numpy(E).cuda())
                return U.cpu().numpy(), S.cpu().numpy(), Vh.cpu().numpy()
        except Exception:
            pass
    return scipy.linalg.svd(E)


class Laplace_QFS(QFS):
    def __init__(self, bdy, interior, slp, dlp, options=None, **kw):
        kw.update(options or {})
        super().__init__(bdy, interior, slp, dlp, LaplaceKernel(), **kw)


class Modified_Helmholtz_QFS(QFS):
    def __init__(self, bdy, interior, slp, dlp, k, options=None, closer_source=True, **kw):
        kw.update(options or {})
        super().__init__(bdy, interior, slp, dlp, ModifiedHelmholtzKernel(k), **kw)


class Stokes_QFS(QFS):
    def __init__(self, bdy, interior, slp, dlp, options=None, **kw):
        kw.update(options or {})
        super().__init__(bdy, interior, slp, dlp, StokesKernel(), **kw)


# ---------------------------------------------------------------------------
def layer_apply_matrix_free(kernel, src, trg_x, trg_y, charge, chunk=2000):
    """Host direct sum  u(x_t) = sum_s G(x_t - y_s) charge_s w_s  (used by setup/tests)."""
    w = np.tile(src.weights, kernel.ncomp)
    q = np.asarray(charge, dtype=float).ravel()*w
    out = np.empty(kernel.ncomp*trg_x.size)
    nt = trg_x.size
    for a in range(0, nt, chunk):
        b = min(nt, a + chunk)
        K = kernel.slp(trg_x[a:b], trg_y[a:b], src.x, src.y)
        res = K.dot(q)
        for c in range(kernel.ncomp):
            out[c*nt + a:c*nt + b] = res[c*(b - a):(c + 1)*(b - a)]
    return out
