"""
ORACLE (test infrastructure, CPU, NumPy/SciPy) -- never imported by the product.

Line-by-line restatement of the reference's inhomogeneous scalar solve:

  ScalarSolver.__call__            ipde/solvers/multi_boundary/scalar.py:72-117
  evaluate_to_grid_pnai            ipde/solvers/multi_boundary/scalar.py:63-71
  ModifiedHelmholtzSolver._grid_solve / _get_specific_operators
                                   ipde/solvers/multi_boundary/modified_helmholtz.py:40-46
  PoissonSolver._grid_solve / _get_specific_operators
                                   ipde/solvers/multi_boundary/poisson.py:30-38
  ScalarHelper.__call__ / correct  ipde/solvers/internals/scalar.py:68-116
  EmbeddedFunction.get_components / load_data
                                   ipde/embedded_function.py:135-138,105-113
  ebdyc.demean_function            ipde/ebdy_collection.py:808-816
  ebdy.interpolate_radial_to_points  ipde/embedded_boundary.py:419-434
  derivatives fd_x_4 / fd_y_4      ipde/derivatives.py:3-23

Third-party arithmetic that is absent from /root/reference is replaced by its
mathematical definition, evaluated directly:
  finufft.nufft2d2 / near_finder periodic_interp2d  ->  direct nonuniform DFT
  pybie2d *_Layer_Apply (backend 'fly')             ->  direct O(S*T) sums, scipy k0
  fast_interp.interp2d(k=3, periodic)               ->  centred 4-point Lagrange stencil
  qfs.*_QFS                                          ->  the dense operators built by
        ipde_b200/qfs.py (host setup), applied here in NumPy (C tau, LU substitution)
Parity status of those pieces: unpinned (no golden vectors exist in the reference);
the annular solver underneath is pinned (oracle/annular.py).

Inputs are plain arrays taken from a registered EmbeddedBoundaryCollection, so the
oracle sees exactly the inputs the CUDA path sees.
"""
import numpy as np
import scipy.special
import scipy.linalg
from .annular import (ApproximateAnnularGeometry, RealAnnularGeometry,
                      AnnularModifiedHelmholtzSolver, AnnularPoissonSolver)

TWO_PI = 2.0*np.pi


def nudft2d_type2(fh, x, y, chunk=512):
    """
    v[f, j] = (1/(Nx Ny)) sum_{k1,k2} fh[f, k1, k2] exp(+i (k1 x_j + k2 y_j)),
    modes in FFT order (finufft modeord=1, isign=+1).  fh: (nf, Nx, Ny) complex.
    """
    fh = np.asarray(fh)
    if fh.ndim == 2:
        fh = fh[None]
    nf, Nx, Ny = fh.shape
    k1 = np.fft.fftfreq(Nx, 1.0/Nx)
    k2 = np.fft.fftfreq(Ny, 1.0/Ny)
    out = np.empty((nf, x.size), dtype=complex)
    for a in range(0, x.size, chunk):
        b = min(x.size, a + chunk)
        A = np.exp(1j*np.outer(x[a:b], k1))
        B = np.exp(1j*np.outer(y[a:b], k2))
        for f in range(nf):
            out[f, a:b] = np.einsum('jk,jk->j', A.dot(fh[f]), B)
    return out/(Nx*Ny)


def fd_x_4(f, h):
    fx = np.zeros_like(f)
    fx[2:-2] = -(f[4:] - 8*f[3:-1] + 8*f[1:-3] - f[:-4])/(12*h)
    return fx


def fd_y_4(f, h):
    fy = np.zeros_like(f)
    fy[:, 2:-2] = -(f[:, 4:] - 8*f[:, 3:-1] + 8*f[:, 1:-3] - f[:, :-4])/(12*h)
    return fy


def lagrange_interp2d_periodic(f, x0, y0, hx, hy, px, py, order=3):
    """Periodic tensor-product Lagrange interpolation with order+1 centred nodes."""
    Nx, Ny = f.shape
    n = order + 1
    def weights(p, lo, h, N):
        s = (p - lo)/h
        i0 = np.floor(s).astype(int) - (n//2 - 1)
        u = s - i0                       # position relative to the first stencil node
        w = np.ones((p.size, n))
        for a in range(n):
            for b in range(n):
                if a != b:
                    w[:, a] *= (u - b)/(a - b)
        idx = (i0[:, None] + np.arange(n)[None, :]) % N
        return idx, w
    ix, wx = weights(px, x0, hx, Nx)
    iy, wy = weights(py, y0, hy, Ny)
    vals = f[ix[:, :, None], iy[:, None, :]]
    return np.einsum('pab,pa,pb->p', vals, wx, wy)


def layer_apply(kind, k, sx, sy, charge_w, tx, ty, chunk=4000):
    """u(t) = sum_s G(|t - s|) charge_w[s];  G = -log r/2pi or K0(k r)/2pi."""
    out = np.empty(tx.size)
    for a in range(0, tx.size, chunk):
        b = min(tx.size, a + chunk)
        dx = tx[a:b, None] - sx[None, :]
        dy = ty[a:b, None] - sy[None, :]
        r2 = dx*dx + dy*dy
        if kind in ('laplace', 'poisson'):
            G = -np.log(r2)/(4*np.pi)
        else:
            G = scipy.special.k0(k*np.sqrt(r2))/TWO_PI
        out[a:b] = G.dot(charge_w)
    return out


def interpolate_radial_to_points(fr, transf_r, t, chunk=2048):
    """Chebyshev-Gauss x Fourier interpolation by a direct NUDFT (embedded_boundary.py:419-434)."""
    M, n = fr.shape
    hold = np.concatenate([fr, fr[::-1]], axis=0)
    kr = np.fft.fftfreq(2*M, 1.0/(2*M))
    funch = np.fft.fft2(hold)*np.exp(-1j*kr[:, None]*np.pi/M/2)
    funch[M] = 0.0
    return nudft2d_type2(funch, transf_r, t, chunk=chunk)[0].real


class OracleScalarHelper(object):
    def __init__(self, ebdy, kind, k, qfs_g, qfs_r):
        self.ebdy, self.kind, self.k = ebdy, kind, k
        self.interior = ebdy.interior
        self.AAG = ApproximateAnnularGeometry(ebdy.bdy.N, ebdy.M, ebdy.radial_width,
                                              ebdy.approximate_radius)
        if kind == 'poisson':
            self.annular_solver = AnnularPoissonSolver(self.AAG)
        else:
            self.annular_solver = AnnularModifiedHelmholtzSolver(self.AAG, k=k)
        bb = ebdy.bdy if self.interior else ebdy.interface
        self.RAG = RealAnnularGeometry(bb.speed, bb.curvature, self.AAG)
        CO = self.AAG.CO
        if self.interior:
            self._bv, self._bn = CO.obc_dirichlet[0], CO.obc_neumann[0]
            self._iv, self._in = CO.ibc_dirichlet[0], CO.ibc_neumann[0]
        else:
            self._bv, self._bn = CO.ibc_dirichlet[0], CO.ibc_neumann[0]
            self._iv, self._in = CO.obc_dirichlet[0], CO.obc_neumann[0]
        self.qfs_g, self.qfs_r = qfs_g, qfs_r

    def apply(self, src, tx, ty, ch):
        return layer_apply(self.kind, self.k, src.x, src.y, ch*src.weights, tx, ty)

    def __call__(self, fr, bv, bx, by, **kwargs):
        ebdy = self.ebdy
        ucn = bx*ebdy.interface.normal_x + by*ebdy.interface.normal_y
        zer = np.zeros_like(bv)
        ur = self.annular_solver.solve(self.RAG, fr, zer, zer, **kwargs)
        self.iterations_last_call = self.annular_solver.iterations_last_call
        urn = self._in.dot(ur)
        slp = urn - ucn
        dlp = bv.copy()
        if not self.interior:
            slp *= -1.0
            dlp *= -1.0
        self.sigma_g = self.qfs_g([slp, dlp])
        self.sigma_r = self.qfs_r([slp, dlp])
        self.ur = ur
        self.slp, self.dlp = slp, dlp
        return self.sigma_g

    def correct(self, ub):
        ebdy = self.ebdy
        w = self.apply(self.qfs_g.source, ebdy.interface.x, ebdy.interface.y, self.sigma_g)
        ub = ub - w
        sigma_r_tot = self.qfs_r.u2s(ub) + self.sigma_r
        rslp = self.apply(self.qfs_r.source, ebdy.radial_x.ravel(), ebdy.radial_y.ravel(), sigma_r_tot)
        self.ur = self.ur + rslp.reshape(self.ur.shape)
        return self.ur


class OracleScalarSolver(object):
    """
    kind: 'poisson' | 'modified_helmholtz';  qfs_pairs: [(qfs_g, qfs_r)] per boundary
    (objects from ipde_b200.qfs, applied with their NumPy __call__/u2s).
    """
    def __init__(self, ebdyc, kind, k=0.0, solver_type='spectral', qfs_pairs=None):
        self.ebdyc, self.kind, self.k, self.solver_type = ebdyc, kind, k, solver_type
        self.helpers = [OracleScalarHelper(e, kind, k, g, r) for e, (g, r) in zip(ebdyc, qfs_pairs)]
        self.kx, self.ky = ebdyc.kx, ebdyc.ky
        self.ikx, self.iky = ebdyc.ikx, ebdyc.iky
        lap = -self.kx*self.kx - self.ky*self.ky
        if kind == 'poisson':
            lap = lap.copy()
            lap[0, 0] = np.inf
            self.mult = 1.0/lap
        else:
            self.mult = 1.0/(k**2 - lap)

    def _grid_solve(self, fc):
        if self.kind == 'poisson':
            e = self.ebdyc
            fc = fc - (np.sum(fc)*e.grid.xh*e.grid.yh)*e.bumpy
        uch = np.fft.fft2(fc)*self.mult
        return uch, np.fft.ifft2(uch).real

    def __call__(self, f, **kwargs):
        e = self.ebdyc
        N = e.grid_phys.N
        f = np.asarray(f)
        fg = np.zeros(e.grid.shape)
        fg[e.phys] = f[:N]
        fc = fg*e.grid_step
        fr_list = [r.copy() for r in e.v2r(f[N:])]
        uch, uc = self._grid_solve(fc)
        self.uc_grid_only = uc.copy()
        if self.solver_type == 'spectral':
            stack = np.stack([uch, self.ikx*uch, self.iky*uch])
            allb = nudft2d_type2(stack, e.interfaces_x_transf, e.interfaces_y_transf)
            bvs, bxs, bys = allb[0].real, allb[1].real, allb[2].real
        else:
            g = e.grid
            I = lambda a: lagrange_interp2d_periodic(a, g.xv[0], g.yv[0], g.xh, g.yh, e.all_ivx, e.all_ivy, 3)
            bvs = I(uc)
            bxs = I(fd_x_4(uc, g.xh))
            bys = I(fd_y_4(uc, g.yh))
        self.interface_values = (bvs, bxs, bys)
        bvl, bxl, byl = e.v2l(bvs), e.v2l(bxs), e.v2l(bys)
        sig = [h(fr, bv, bx, by, **kwargs) for h, fr, bv, bx, by in zip(self.helpers, fr_list, bvl, bxl, byl)]
        self.iteration_counts = [h.iterations_last_call for h in self.helpers]
        # all grid-side sources -> (phys not in annulus) U (all interface points)
        sx = np.concatenate([h.qfs_g.source.x for h in self.helpers])
        sy = np.concatenate([h.qfs_g.source.y for h in self.helpers])
        sw = np.concatenate([h.qfs_g.source.weights for h in self.helpers])
        out = layer_apply(self.kind, self.k, sx, sy, np.concatenate(sig)*sw, e.grid_pnai.x, e.grid_pnai.y)
        gu, bus = e.divide_pnai(out)
        uc[e.phys_not_in_annulus] += gu
        urs = [h.correct(bu) for h, bu in zip(self.helpers, bus)]
        for ebdy, ur in zip(e, urs):
            uc[ebdy.grid_ia_xind, ebdy.grid_ia_yind] = \
                interpolate_radial_to_points(ur, ebdy.grid_ia_r_transf, ebdy.grid_ia_t)
        uc *= e.phys
        return np.concatenate([uc[e.phys]] + [ur.ravel() for ur in urs])
