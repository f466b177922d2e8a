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
    # boundaries with more nodes than this are factored on the device (needs CUDA)
    DEVICE_BUILD_MIN_N = 1500

    def __init__(self, bdy, interior, slp, dlp, kernel, tol=1e-12,
                 source_upsample_factor=1.0, alpha_s=None, alpha_c=None, rho=None, rho_f=None, u2s=True,
                 device_build=None, operators=True, b2c=True):
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
        self.want_u2s = u2s
        self.want_b2c = b2c                  # False: only the u2s operators (homogeneous correction)
        self.want_operators = operators      # False: source/check geometry only (boundaries owned by another rank)
        self.on_device = False
        if device_build is None:
            device_build = bdy.N > self.DEVICE_BUILD_MIN_N and _cuda_available()
        self._build(device_build)

    def _build(self, device_build=False):
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
        if not self.want_operators:
            self.A = self.Sinv = self.V = self.Ab = self.Sinv_b = self.Vb = None
            return
        if device_build:
            return self._build_stokes_on_device(bM) if nc == 2 else self._build_on_device(bM)
        w_s = np.tile(src.weights, nc)
        UM = fourier_resample_matrix(N, M)
        U_MN = UM if nc == 1 else scipy.linalg.block_diag(*([UM]*nc))
        # check <- boundary densities through the upsampled trapezoid rule
        Nf = self.rho_f*N
        bF = GSB(c=fourier_resample(bdy.c, Nf))
        UF = fourier_resample_matrix(N, Nf)
        w_f = bF.weights
        if self.want_b2c:
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
        else:
            self.A = self.Sinv = self.V = None
        if self.want_u2s:
            Eb = ker.slp(bM.x, bM.y, src.x, src.y)*w_s[None, :]
            self.Ab, self.Sinv_b, self.Vb = self._factor(Eb, U_MN)
        else:
            self.Ab = self.Sinv_b = self.Vb = None

    def _build_on_device(self, bM):
        """
        The same operators, built on the GPU (setup only): kernel matrices by ipde_kernel_matrix (our K0/K1),
        the product with the Fourier upsampling matrix as row FFTs (our Stockham FFT), SVD and the U^T C
        products by torch.linalg (cuSOLVER / cuBLAS -- setup-time library calls, not on the per-solve path).
        A, Sinv, V (and the u2s set) are left on the device as torch tensors.
        """
        import torch
        from .device import to_dev, empty, ptr, stream, ensure_k0_table, FFT1D
        from ._lib import lib
        L = lib()
        bdy, ker, src, chk = self.bdy, self.kernel, self.source, self.check
        kind = 0 if ker.name == 'laplace' else 1
        if kind == 1:
            ensure_k0_table()
        N, M = self.N, self.M
        Nf = self.rho_f*N
        bF = GSB(c=fourier_resample(bdy.c, Nf))
        d = lambda a: to_dev(np.ascontiguousarray(a, dtype=float))
        cx, cy = d(chk.x), d(chk.y)
        fx, fy, fnx, fny, fw = d(bF.x), d(bF.y), d(bF.normal_x), d(bF.normal_y), d(bF.weights)
        sx, sy, sw = d(src.x), d(src.y), d(src.weights)
        fft_f, fft_c = FFT1D(Nf), FFT1D(N)
        # coarse mode k' <- fine mode index (fourier_resample, upsampling branch)
        hh = (N - 1)//2
        idx = np.zeros(N, dtype=np.int64)
        idx[:hh + 1] = np.arange(hh + 1)
        idx[N - hh:] = Nf - hh + np.arange(hh)
        d_idx = to_dev(idx)
        layers = [l for l, on in ((0, self.has_slp), (1, self.has_dlp)) if on]
        if not self.want_b2c:
            layers = []
        C = empty((M, max(1, len(layers))*N))
        rows = max(1, min(M, int(2.0e9//(16*Nf))))
        K = empty((rows, Nf))
        for li, layer in enumerate(layers):
            for r0 in range(0, M, rows):
                r1 = min(M, r0 + rows)
                nr = r1 - r0
                L.ipde_kernel_matrix(kind, float(ker.k), layer, ptr(cx[r0:r1]), ptr(cy[r0:r1]), nr, ptr(fx), ptr(fy),
                                     ptr(fnx), ptr(fny), ptr(fw), Nf, ptr(K), stream())
                Kc = K[:nr].to(torch.complex128)
                R = fft_f.c2c(Kc, +1)                       # sum_j r_j exp(+2 pi i j k / Nf)
                Z = R[:, d_idx].contiguous()
                if N % 2 == 0:
                    q = N//2
                    Z[:, q] = 0.5*(R[:, q] + R[:, Nf - q])
                out = fft_c.c2c(Z, -1, 1.0/N)
                C[r0:r1, li*N:(li + 1)*N] = out.real
                del Kc, R, Z, out
        del K
        if self.want_b2c:
            E = empty((M, M))
            L.ipde_kernel_matrix(kind, float(ker.k), 0, ptr(cx), ptr(cy), M, ptr(sx), ptr(sy), 0, 0, ptr(sw), M, ptr(E),
                                 stream())
            self.A, self.Sinv, self.V = self._factor_device(E, C)
            del E
        else:
            self.A = self.Sinv = self.V = None
        del C
        if self.want_u2s:
            Eb = empty((M, M))
            bx, by = d(bM.x), d(bM.y)
            L.ipde_kernel_matrix(kind, float(ker.k), 0, ptr(bx), ptr(by), M, ptr(sx), ptr(sy), 0, 0, ptr(sw), M, ptr(Eb),
                                 stream())
            UM = to_dev(fourier_resample_matrix(N, M))
            self.Ab, self.Sinv_b, self.Vb = self._factor_device(Eb, UM)
            del Eb, UM
        else:
            self.Ab = self.Sinv_b = self.Vb = None
        self.on_device = True

    def _build_stokes_on_device(self, bM):
        """
        Stokes operators of a large boundary (setup only): the 2x2 block kernel matrices are formed with torch
        elementwise ops on the GPU in row chunks, multiplied by the Fourier upsampling matrix and factored there
        (torch.linalg -- setup-time library calls).  Same matrices as the host branch of _build.
        """
        import torch
        from .device import to_dev
        bdy, src, chk = self.bdy, self.source, self.check
        N, M = self.N, self.M
        Nf = self.rho_f*N
        bF = GSB(c=fourier_resample(bdy.c, Nf))
        d = lambda a: to_dev(np.ascontiguousarray(a, dtype=float))
        UF = d(fourier_resample_matrix(N, Nf))                 # (Nf, N)
        fx, fy, fnx, fny, fw = d(bF.x), d(bF.y), d(bF.normal_x), d(bF.normal_y), d(bF.weights)
        cx, cy = d(chk.x), d(chk.y)
        c4 = 1.0/(4.0*np.pi)

        def stokeslet(tx, ty, sx, sy, w):
            dx, dy = tx[:, None] - sx[None, :], ty[:, None] - sy[None, :]
            r2 = dx*dx + dy*dy
            Lg = -0.5*torch.log(r2)
            xy = c4*dx*dy/r2*w
            return (c4*(Lg + dx*dx/r2)*w, xy, xy, c4*(Lg + dy*dy/r2)*w)      # xx, xy, yx, yy

        def stresslet(tx, ty, sx, sy, nx, ny, w):
            dx, dy = tx[:, None] - sx[None, :], ty[:, None] - sy[None, :]
            r2 = dx*dx + dy*dy
            f = (dx*nx[None, :] + dy*ny[None, :])/(r2*r2)/np.pi*w
            xy = f*dx*dy
            return (f*dx*dx, xy, xy, f*dy*dy)

        layers = [l for l, on in ((0, self.has_slp), (1, self.has_dlp)) if on]
        C = torch.empty((2*M, len(layers)*2*N), dtype=torch.float64, device=UF.device)
        rows = max(1, min(M, int(1.0e8//Nf)))
        for li, layer in enumerate(layers):
            for r0 in range(0, M, rows):
                r1 = min(M, r0 + rows)
                if layer == 0:
                    blocks = stokeslet(cx[r0:r1], cy[r0:r1], fx, fy, fw[None, :])
                else:
                    blocks = stresslet(cx[r0:r1], cy[r0:r1], fx, fy, fnx, fny, fw[None, :])
                for a in range(2):                      # target component
                    for i in range(2):                  # source component
                        C[a*M + r0:a*M + r1, li*2*N + i*N:li*2*N + (i + 1)*N] = blocks[2*a + i] @ UF
                del blocks
        sx, sy, sw = d(src.x), d(src.y), d(src.weights)

        def slp_matrix(tx, ty):
            xx, xy, yx, yy = stokeslet(tx, ty, sx, sy, sw[None, :])
            return torch.cat([torch.cat([xx, xy], dim=1), torch.cat([yx, yy], dim=1)], dim=0)

        self.A, self.Sinv, self.V = self._factor_device(slp_matrix(cx, cy), C)
        del C
        if self.want_u2s:
            UM = d(fourier_resample_matrix(N, M))
            Z = torch.zeros_like(UM)
            U_MN = torch.cat([torch.cat([UM, Z], dim=1), torch.cat([Z, UM], dim=1)], dim=0)
            self.Ab, self.Sinv_b, self.Vb = self._factor_device(slp_matrix(d(bM.x), d(bM.y)), U_MN)
        else:
            self.Ab = self.Sinv_b = self.Vb = None
        self.on_device = True

    def _factor_device(self, E, rhs_op):
        import torch
        U, S, Vh = torch.linalg.svd(E, full_matrices=False, driver='gesvd')
        keep = S > self.svd_cut*S[0]
        if int(keep.sum().item()) % 2:          # even count: rows stay 16-byte aligned for the device GEMV
            keep[torch.nonzero(keep)[-1, 0]] = False
        A = (U.T[keep] @ rhs_op).contiguous()
        return A, (1.0/S[keep]).contiguous(), Vh[keep].T.contiguous()

    def _factor(self, E, rhs_op):
        """E = U S V^T  ->  (U^T rhs_op, 1/S with tiny singular values dropped, V), all C-contiguous."""
        U, S, Vt = _svd(E)
        keep = S > self.svd_cut*S[0]
        if int(keep.sum()) % 2:                 # even count: rows stay 16-byte aligned for the device GEMV
            keep[np.nonzero(keep)[0][-1]] = False
        # truncated directions contribute nothing: drop their rows of U^T rhs and columns of V (~25% of the
        # per-solve operator traffic)
        return (np.ascontiguousarray(U.T[keep].dot(rhs_op)), np.ascontiguousarray(1.0/S[keep]),
                np.ascontiguousarray(Vt[keep].T))

    @staticmethod
    def _times_upsample(K, UF, nc):
        Nf, N = UF.shape
        return np.hstack([K[:, i*Nf:(i + 1)*Nf].dot(UF) for i in range(nc)])

    def __call__(self, tau_list):
        tau = np.concatenate([np.asarray(t, dtype=float).ravel() for t in tau_list])
        if self.on_device:
            import torch
            t = torch.from_numpy(tau).to(self.A.device)
            return (self.V @ (self.Sinv*(self.A @ t))).cpu().numpy()
        return self.V.dot(self.Sinv*self.A.dot(tau))

    def u2s(self, ub):
        if self.Ab is None:
            raise Exception('this QFS object was built without the u2s operators (u2s=False)')
        ub = np.asarray(ub, dtype=float).ravel()
        if self.on_device:
            import torch
            t = torch.from_numpy(ub).to(self.Ab.device)
            return (self.Vb @ (self.Sinv_b*(self.Ab @ t))).cpu().numpy()
        return self.Vb.dot(self.Sinv_b*self.Ab.dot(ub))


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _svd(E):
    """SVD on the host (small systems and the CPU oracle); large boundaries use QFS._build_on_device."""
    return scipy.linalg.svd(E)


# boundaries with at least this many nodes get the structured (Fourier-banded) operators of qfs_structured.py;
# IPDE_QFS=dense / IPDE_QFS=structured overrides, as does the `structured=` keyword
STRUCTURED_MIN_N = 1024
STRUCTURED_STOKES = True


def _make_qfs(kernel, bdy, interior, slp, dlp, options=None, structured=None, **kw):
    """
    Factory behind the reference's constructor names (qfs.Laplace_QFS(bdy, interior, slp, dlp, options), ...):
    dense SVD-factored operators for small boundaries (and whenever `device_build` is given explicitly), structured
    O(n log n) operators for large ones.  Both expose `source`, `__call__([slp, dlp])`, `u2s(ub)`, `N`, `M`.
    """
    import os
    kw.update(options or {})
    explicit = structured is not None
    if structured is None:
        env = os.environ.get('IPDE_QFS', '')
        if env == 'structured':
            structured = True
        elif env == 'dense' or 'device_build' in kw:
            structured = False
        else:
            structured = bdy.N >= STRUCTURED_MIN_N and (kernel.ncomp == 1 or STRUCTURED_STOKES)
    if structured:
        from .qfs_structured import StructuredQFS, NotResolved
        skw = {k: v for k, v in kw.items() if k != 'device_build'}
        try:
            return StructuredQFS(bdy, interior, slp, dlp, kernel, **skw)
        except NotResolved:
            if explicit:
                raise
    return QFS(bdy, interior, slp, dlp, kernel, **kw)


def Laplace_QFS(bdy, interior, slp, dlp, options=None, **kw):
    return _make_qfs(LaplaceKernel(), bdy, interior, slp, dlp, options, **kw)


def Modified_Helmholtz_QFS(bdy, interior, slp, dlp, k, options=None, closer_source=True, **kw):
    return _make_qfs(ModifiedHelmholtzKernel(k), bdy, interior, slp, dlp, options, **kw)


def Stokes_QFS(bdy, interior, slp, dlp, options=None, **kw):
    return _make_qfs(StokesKernel(), bdy, interior, slp, dlp, options, **kw)


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
