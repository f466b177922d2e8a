"""
ORACLE (test infrastructure, CPU, NumPy/SciPy) -- never imported by the product.

O(N^2 log N) CPU statement of the scalar inhomogeneous solve, for problem sizes where the direct NUDFT and the
all-pairs sums of oracle/solvers.py (the reference's default 'pybie2d' backend) stop being feasible.  Same algorithm,
same call order as oracle/solvers.py (ScalarSolver.__call__, ipde/solvers/multi_boundary/scalar.py:72-117); only the
two evaluation primitives are replaced by their fast equivalents, as the reference itself offers
(finufft for the interpolation, scalar.py:84-88; FMM / grid evaluator for the sums, modified_helmholtz.py:48-67):

  nufft2d2             type-2 NUFFT: sigma = 2 oversampled inverse FFT + exponential-of-semicircle kernel, w = 16
                       (the algorithm of finufft.nufft2d2, which the reference calls at eps = 1e-14)
  ewald_sum            free-space sum of G over all sources at grid nodes and off-grid points by spectral Ewald
                       (Gaussian split; truncated-kernel FFT solve on a box of period >= extent + L), the CPU
                       counterpart of the reference's ScalarFreespaceGridEvaluator
                       (ipde/grid_evaluators/scalar_grid_evaluator.py:188-231,266-307) without its 1e-10 ceiling

Pinned by tests/test_oracle.py against the direct versions (NUDFT, all-pairs sums) at small sizes.
FFTs go through scipy.fft with `workers` threads (the "threaded" CPU line); workers=1 and numpy.fft
give the reference-faithful single-thread line.
"""
import numpy as np
import scipy.fft
import scipy.special as ss
from scipy.spatial import cKDTree

from ipde_b200.ewald import EwaldParameters, trunc_kernel_hat, local_tables, eval_local
from .solvers import OracleScalarSolver, OracleScalarHelper

TWO_PI = 2.0*np.pi


# ------------------------------------------------------------------------------------------------------------------
def _es_setup(w=16):
    beta = 2.30*w
    z, wq = np.polynomial.legendre.leggauss(200)
    phi = np.exp(beta*(np.sqrt(1.0 - z*z) - 1.0))
    return beta, z, wq, phi


def _es_deconv(n, nf, w, beta, z, wq, phi):
    """1/phihat at the n coarse modes (FFT order) for an nf-point fine grid"""
    k = np.fft.fftfreq(n, 1.0/n)
    xi = np.pi*w*np.abs(k)/nf
    I = (np.cos(np.outer(xi, z))*phi).dot(wq)
    return 2.0/(w*I)


def nufft2d2(fh, x, y, workers=1, w=16, sigma=2):
    """
    v[f, j] = (1/(Nx Ny)) sum_k fh[f, k1, k2] exp(+i (k1 x_j + k2 y_j)), modes in FFT order; x, y in [0, 2 pi).
    Same definition as oracle.solvers.nudft2d_type2 (finufft modeord=1, isign=+1, divided by Nx Ny).
    """
    fh = np.asarray(fh)
    if fh.ndim == 2:
        fh = fh[None]
    nf_, Nx, Ny = fh.shape
    fx, fy = sigma*Nx, sigma*Ny
    beta, z, wq, phi = _es_setup(w)
    dx = _es_deconv(Nx, fx, w, beta, z, wq, phi)
    dy = _es_deconv(Ny, fy, w, beta, z, wq, phi)
    kx = np.fft.fftfreq(Nx, 1.0/Nx).astype(int) % fx
    ky = np.fft.fftfreq(Ny, 1.0/Ny).astype(int) % fy
    # fine-grid positions and kernel weights of every target
    def stencil(p, n):
        s = np.mod(p, TWO_PI)*n/TWO_PI
        i0 = np.ceil(s - w/2.0).astype(np.int64)
        idx = i0[:, None] + np.arange(w)[None, :]
        zz = (idx - s[:, None])*(2.0/w)
        wt = np.where(np.abs(zz) < 1.0, np.exp(beta*(np.sqrt(np.clip(1.0 - zz*zz, 0.0, None)) - 1.0)), 0.0)
        return idx % n, wt
    ix, wx = stencil(x, fx)
    iy, wy = stencil(y, fy)
    out = np.empty((nf_, x.size), dtype=complex)
    for f in range(nf_):
        pad = np.zeros((fx, fy), dtype=complex)
        pad[np.ix_(kx, ky)] = fh[f]*dx[:, None]*dy[None, :]
        fine = scipy.fft.ifft2(pad, workers=workers)*(fx*fy)
        vals = fine[ix[:, :, None], iy[:, None, :]]
        out[f] = np.einsum('pab,pa,pb->p', vals, wx, wy)
    return out/(Nx*Ny)


# ------------------------------------------------------------------------------------------------------------------
class EwaldSum(object):
    """
    u(t) = sum_s G(|t - y_s|) q_s, G = -log r/2pi (kind 0) or K0(kappa r)/2pi (kind 1), for fixed sources and targets:
    grid nodes (ix, iy) of the background grid (x0 + h ix, y0 + h iy) and/or off-grid points.
    """
    def __init__(self, kind, kappa, x0, y0, h, sx, sy, grid_nodes=None, points=None, workers=1):
        self.kind, self.kappa, self.h, self.workers = kind, kappa, h, workers
        P = EwaldParameters(h)
        self.P = P
        self.sx, self.sy = np.asarray(sx, dtype=float), np.asarray(sy, dtype=float)
        xs, ys = [self.sx], [self.sy]
        if grid_nodes is not None:
            self.gix, self.giy = np.asarray(grid_nodes[0]), np.asarray(grid_nodes[1])
            xs.append(x0 + h*self.gix); ys.append(y0 + h*self.giy)
        if points is not None:
            self.px, self.py = np.asarray(points[0], dtype=float), np.asarray(points[1], dtype=float)
            xs.append(self.px); ys.append(self.py)
        self.has_grid, self.has_pts = grid_nodes is not None, points is not None
        lox, hix = min(a.min() for a in xs), max(a.max() for a in xs)
        loy, hiy = min(a.min() for a in ys), max(a.max() for a in ys)
        D = np.hypot(hix - lox, hiy - loy)
        m = P.m
        self.L = D + 2.0*np.sqrt(2.0)*m*h + 2.0*h
        # box aligned with the background grid, period >= extent + L
        self.ilo_x = int(np.floor((lox - x0)/h)) - m - 1
        self.ilo_y = int(np.floor((loy - y0)/h)) - m - 1
        nx = int(np.ceil((hix - lox)/h)) + 2*m + 3 + int(np.ceil(self.L/h))
        ny = int(np.ceil((hiy - loy)/h)) + 2*m + 3 + int(np.ceil(self.L/h))
        self.nx, self.ny = scipy.fft.next_fast_len(nx, real=True), scipy.fft.next_fast_len(ny, real=True)
        self.bx0, self.by0 = x0 + h*self.ilo_x, y0 + h*self.ilo_y
        kx = np.fft.fftfreq(self.nx, h/TWO_PI)[:, None]
        ky = np.fft.rfftfreq(self.ny, h/TWO_PI)[None, :]
        K2 = kx*kx + ky*ky
        GT = trunc_kernel_hat(kind, kappa, self.L, np.sqrt(K2))
        self.multH = GT
        self.multU = GT*np.exp(-K2/(8*P.xi**2))
        self.ck = np.exp(-kappa*kappa/(4*P.xi**2)) if kind == 1 else 1.0
        self.tab = local_tables(kind, kappa, P.xi, P.rc)
        # spreading stencils of the sources
        self.six, self.siy, self.swt = self._window(self.sx, self.sy)
        tree = cKDTree(np.column_stack([self.sx, self.sy]))
        if self.has_grid:
            tx, ty = x0 + h*self.gix, y0 + h*self.giy
            self.g_pairs = self._pairs(tree, tx, ty)
        if self.has_pts:
            self.p_pairs = self._pairs(tree, self.px, self.py)
            self.pix, self.piy, self.pwt = self._window(self.px, self.py)

    def _window(self, px, py):
        """(index arrays, separable Gaussian window weights) of the (2m+1)^2 box nodes around every point"""
        P, h, m = self.P, self.h, self.P.m
        cx = np.rint((px - self.bx0)/h).astype(np.int64)
        cy = np.rint((py - self.by0)/h).astype(np.int64)
        off = np.arange(-m, m + 1)
        ix = cx[:, None] + off[None, :]
        iy = cy[:, None] + off[None, :]
        gx = np.exp(-2*P.xi**2*(self.bx0 + h*ix - px[:, None])**2)
        gy = np.exp(-2*P.xi**2*(self.by0 + h*iy - py[:, None])**2)
        return ix, iy, (gx, gy)

    def _pairs(self, tree, tx, ty):
        """(target index, source index, r^2) of all pairs closer than rc"""
        tt = cKDTree(np.column_stack([tx, ty]))
        sp = tt.sparse_distance_matrix(tree, self.P.rc, output_type='coo_matrix')
        return sp.row, sp.col, sp.data**2, tx.size

    def _local(self, pairs, q):
        ti, si, r2, nt = pairs
        return np.bincount(ti, weights=eval_local(self.tab, r2)*q[si], minlength=nt)

    def __call__(self, q):
        """returns (values at the grid nodes or None, values at the points or None)"""
        P, h = self.P, self.h
        q = np.asarray(q, dtype=float)
        c = 2*P.xi**2/np.pi*self.ck
        gx, gy = self.swt
        S = np.zeros((self.nx, self.ny))
        contrib = (c*q)[:, None, None]*gx[:, :, None]*gy[:, None, :]
        np.add.at(S, (self.six[:, :, None], self.siy[:, None, :]), contrib)
        Sh = scipy.fft.rfft2(S, workers=self.workers)
        out_g = out_p = None
        if self.has_grid:
            U = scipy.fft.irfft2(Sh*self.multU, s=S.shape, workers=self.workers)
            out_g = U[self.gix - self.ilo_x, self.giy - self.ilo_y] + self._local(self.g_pairs, q)
        if self.has_pts:
            H = scipy.fft.irfft2(Sh*self.multH, s=S.shape, workers=self.workers)
            wx, wy = self.pwt
            vals = H[self.pix[:, :, None], self.piy[:, None, :]]
            far = h*h*(2*P.xi**2/np.pi)*np.einsum('pab,pa,pb->p', vals, wx, wy)
            out_p = far + self._local(self.p_pairs, q)
        return out_g, out_p


# ------------------------------------------------------------------------------------------------------------------
class FastScalarHelper(OracleScalarHelper):
    """OracleScalarHelper with the two sums of correct() (internals/scalar.py:101,113) done by EwaldSum"""
    def set_grid(self, grid, workers):
        e = self.ebdy
        kind = 0 if self.kind == 'poisson' else 1
        g, r = self.qfs_g.source, self.qfs_r.source
        x0, y0, h = grid.xv[0], grid.yv[0], grid.xh
        self.ev_w = EwaldSum(kind, self.k, x0, y0, h, g.x, g.y, points=(e.interface.x, e.interface.y), workers=workers)
        self.ev_r = EwaldSum(kind, self.k, x0, y0, h, r.x, r.y, points=(e.radial_x.ravel(), e.radial_y.ravel()),
                             workers=workers)

    def correct(self, ub):
        g, r = self.qfs_g.source, self.qfs_r.source
        w = self.ev_w(self.sigma_g*g.weights)[1]
        sigma_r_tot = self.qfs_r.u2s(ub - w) + self.sigma_r
        rslp = self.ev_r(sigma_r_tot*r.weights)[1]
        self.ur = self.ur + rslp.reshape(self.ur.shape)
        return self.ur


class FastScalarSolver(OracleScalarSolver):
    """
    oracle.solvers.OracleScalarSolver with FFT-based evaluation primitives (module docstring).  `workers`: threads of
    scipy.fft; `fft='numpy'` uses numpy.fft complex transforms for the grid solve exactly as the reference falls back
    to (ipde/utilities.py:9-12).
    """
    def __init__(self, ebdyc, kind, k=0.0, qfs_pairs=None, workers=1, fft='scipy'):
        super().__init__(ebdyc, kind, k, 'spectral', qfs_pairs)
        self.workers, self.fft = workers, fft
        self.helpers = [FastScalarHelper(e, kind, k, g, r) for e, (g, r) in zip(ebdyc, qfs_pairs)]
        e = ebdyc
        for h in self.helpers:
            h.set_grid(e.grid, workers)
        sx = np.concatenate([h.qfs_g.source.x for h in self.helpers])
        sy = np.concatenate([h.qfs_g.source.y for h in self.helpers])
        self.sw = np.concatenate([h.qfs_g.source.weights for h in self.helpers])
        pix, piy = np.nonzero(e.phys_not_in_annulus)
        self.ev_g = EwaldSum(0 if kind == 'poisson' else 1, k, e.grid.xv[0], e.grid.yv[0], e.grid.xh, sx, sy,
                             grid_nodes=(pix, piy), points=(e.all_ivx, e.all_ivy), workers=workers)

    def _grid_solve(self, fc):
        if self.kind == 'poisson':
            e = self.ebdyc
            fc = fc - (np.sum(fc)*e.grid.xh*e.grid.yh)*e.bumpy
        if self.fft == 'numpy':
            uch = np.fft.fft2(fc)*self.mult
            return uch, np.fft.ifft2(uch).real
        uch = scipy.fft.fft2(fc, workers=self.workers)*self.mult
        return uch, scipy.fft.ifft2(uch, workers=self.workers).real

    def __call__(self, f, **kwargs):
        e = self.ebdyc
        N = e.grid_phys.N
        f = np.asarray(f)
        fg = np.zeros(e.grid.shape)
        fg[e.phys] = f[:N]
        fc = fg*e.grid_step
        fr_list = [r.copy() for r in e.v2r(f[N:])]
        uch, uc = self._grid_solve(fc)
        stack = np.stack([uch, self.ikx*uch, self.iky*uch])
        allb = nufft2d2(stack, e.interfaces_x_transf, e.interfaces_y_transf, workers=self.workers)
        bvl, bxl, byl = e.v2l(allb[0].real), e.v2l(allb[1].real), e.v2l(allb[2].real)
        sig = [h(fr, bv, bx, by, **kwargs) for h, fr, bv, bx, by in zip(self.helpers, fr_list, bvl, bxl, byl)]
        self.iteration_counts = [h.iterations_last_call for h in self.helpers]
        gu, bu_all = self.ev_g(np.concatenate(sig)*self.sw)
        uc[e.phys_not_in_annulus] += gu
        urs = [h.correct(bu) for h, bu in zip(self.helpers, e.v2l(bu_all))]
        for ebdy, ur in zip(e, urs):
            uc[ebdy.grid_ia_xind, ebdy.grid_ia_yind] = radial_to_points_fast(ur, ebdy.grid_ia_r_transf, ebdy.grid_ia_t,
                                                                               self.workers)
        uc *= e.phys
        return np.concatenate([uc[e.phys]] + [ur.ravel() for ur in urs])


def radial_to_points_fast(fr, transf_r, t, workers=1):
    """ebdy.interpolate_radial_to_points (ipde/embedded_boundary.py:419-434) with the type-2 NUFFT above"""
    M, n = fr.shape
    hold = np.concatenate([fr, fr[::-1]], axis=0)
    kr = np.fft.fftfreq(2*M, 1.0/(2*M))
    funch = np.fft.fft2(hold)*np.exp(-1j*kr[:, None]*np.pi/M/2)
    funch[M] = 0.0
    return nufft2d2(funch, transf_r, t, workers=workers)[0].real
