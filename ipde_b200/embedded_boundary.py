"""
EmbeddedBoundary: one smooth boundary, its offset *interface* curve, the M x N
radial (Chebyshev-Gauss x periodic) grid between them, and -- once a background
grid is registered -- the list of grid points inside that annulus with their
(r, t) coordinates and cutoff values.

API mirror of the reference class (ipde/embedded_boundary.py:55-557): same
constructor `EmbeddedBoundary(bdy, interior, M, h, **kwargs)` (:56-144), same
attribute names (`radial_x/y/rv/shape/speed`, `interface`, `grid_ia_xind/yind/r/t`,
`grid_ia_r_transf`, `grid_to_radial_step`, `radial_cutoff`, `approximate_radius`,
`radial_targ`, `D00`, ...), same helpers (`convert_uv_to_rt`, `nufft_transform_r`,
`interpolate_radial_to_boundary*`, `save`/`regenerate`).  All of this is setup
(host NumPy).  The one per-solve member, `interpolate_radial_to_points`
(:419-434), is served by the CUDA radial->grid kernel when called through the
solvers; the method kept here for drop-in use calls that same kernel.
"""
import numpy as np
from .geometry import GSB, PointSet
from .near import gridpoints_near_curve
from .slepian import SlepianMollifier
from .utilities import affine_transformation, get_chebyshev_nodes, SimpleFourierFilter


def setit(name, dictionary, default):
    return dictionary[name] if name in dictionary else default


def fejer_1(n):
    """Fejer-1 weights on the Chebyshev-Gauss nodes (embedded_boundary.py:21-36)."""
    points = -np.cos(np.pi*(np.arange(n) + 0.5)/n)
    Nodd = np.arange(1, n, 2)
    length = len(Nodd)
    m = n - length
    K = np.arange(m)
    v0 = np.concatenate([2*np.exp(1j*np.pi*K/n)/(1 - 4*K**2), np.zeros(length + 1)])
    v1 = v0[:-1] + np.conjugate(v0[:0:-1])
    w = np.fft.ifft(v1)
    return points, w.real


def LoadEmbeddedBoundary(d):
    bdy = GSB(x=d['bx'], y=d['by'])
    return EmbeddedBoundary(bdy, d['interior'], d['M'], d['h'], **d['kwargs'])


def smoothit(f, factor):
    fh = np.fft.rfft(f)
    n = f.size
    k = np.fft.rfftfreq(n, 1/n)
    max_k = np.abs(k).max()
    fh *= np.exp(-100*(np.abs(k)/max_k)**16)
    return np.fft.irfft(fh, n)


class EmbeddedBoundary(object):
    def __init__(self, bdy, interior, M, h, **kwargs):
        self.bdy = bdy
        self.interior = interior
        self.M = M
        self.h = h
        self.pad_zone             = setit('pad_zone',             kwargs, 0)
        self.coordinate_tolerance = setit('coordinate_tolerance', kwargs, 1e-10)
        self.qfs_tolerance        = setit('qfs_tolerance',        kwargs, 1e-10)
        self.qfs_fsuf             = setit('qfs_fsuf',             kwargs, None)
        self.qfs_FF               = setit('qfs_FF',               kwargs, 0.0)
        self.coordinate_scheme    = setit('coordinate_scheme',    kwargs, 'nufft')
        self.smoothing_factor     = setit('smoothing_factor',     kwargs, None)
        if 'grid_to_radial_step' in kwargs and 'radial_cutoff' in kwargs:
            self.grid_to_radial_step = kwargs['grid_to_radial_step']
            self.radial_cutoff = kwargs['radial_cutoff']
            self.have_heavisides = True
        else:
            self.heaviside = kwargs['heaviside'] if 'heaviside' in kwargs \
                else SlepianMollifier(2*self.M).step
            self.have_heavisides = False
        self.radial_width = self.M*self.h
        self.heaviside_width = self.radial_width - self.pad_zone*self.h
        if self.smoothing_factor is not None:
            self.bdy = GSB(smoothit(self.bdy.x, self.smoothing_factor),
                           smoothit(self.bdy.y, self.smoothing_factor))
        self._generate_radial_grid()
        self.near_radial_guess_inds = \
            np.tile(np.arange(self.bdy.N), self.M).reshape(self.M, self.bdy.N)
        self.kwargs = kwargs

    def regenerate(self, bx, by):
        new_kwargs = self.kwargs.copy()
        new_kwargs.pop('grid_to_radial_step', None)
        new_kwargs.pop('radial_cutoff', None)
        return EmbeddedBoundary(GSB(x=bx.copy(), y=by.copy()), self.interior,
                                self.M, self.h, **new_kwargs)

    def save(self):
        short_kwargs = {k: v for k, v in self.kwargs.items() if k != 'heaviside'}
        return {'bx': self.bdy.x, 'by': self.bdy.y, 'interior': self.interior,
                'M': self.M, 'h': self.h, 'kwargs': short_kwargs}

    ############################################################################
    def register_grid(self, grid, close=None, int_helper1=None, int_helper2=None,
                      float_helper=None, bool_helper=None, index=None,
                      danger_zone_distance=None, verbose=False):
        """
        Locate the grid points in this annulus and their (r, t) coordinates
        (embedded_boundary.py:185-269).  The scratch-array arguments of the
        reference signature are accepted and ignored.
        """
        self.grid = grid
        ddd = danger_zone_distance if danger_zone_distance is not None else 0.0
        out = gridpoints_near_curve(self.bdy, grid.xv, grid.yv,
                                    self.radial_width + ddd,
                                    tol=min(self.coordinate_tolerance, 1e-13))
        self.near_curve_result = out
        iax, iay, r, t = out.ix, out.iy, out.r, out.t
        ia_good, _, _ = self.check_if_r_in_annulus(r)
        self.grid_ia_xind = iax[ia_good]
        self.grid_ia_yind = iay[ia_good]
        self.grid_ia_x = grid.xv[self.grid_ia_xind]
        self.grid_ia_y = grid.yv[self.grid_ia_yind]
        self.grid_ia_r = r[ia_good]
        self.grid_ia_t = t[ia_good]
        self.grid_ia_r_transf = self.nufft_transform_r(self.grid_ia_r)
        self.interface_x_transf = affine_transformation(
            self.interface.x, grid.x_bounds[0], grid.x_bounds[1], 0.0, 2*np.pi)
        self.interface_y_transf = affine_transformation(
            self.interface.y, grid.y_bounds[0], grid.y_bounds[1], 0.0, 2*np.pi)
        if not self.have_heavisides:
            lb = -self.heaviside_width if self.interior else self.heaviside_width
            grts = affine_transformation(self.grid_ia_r, lb, 0, -1, 1)
            self.grid_to_radial_step = 1.0 - self.heaviside(grts)
            arts = affine_transformation(self.radial_rv, lb, 0, -1, 1)
            self.radial_cutoff = self.heaviside(arts)
            self.kwargs['grid_to_radial_step'] = self.grid_to_radial_step
            self.kwargs['radial_cutoff'] = self.radial_cutoff
        if danger_zone_distance is not None:
            idz = np.logical_and(r <= ddd, r >= -self.radial_width - ddd) if self.interior \
                else np.logical_and(r >= -ddd, r <= self.radial_width + ddd)
            self.grid_in_danger_zone_x = iax[idz]
            self.grid_in_danger_zone_y = iay[idz]
            self.grid_in_danger_zone_gi = (t[idz]//self.bdy.dt).astype(int)

    def register_ia_inds(self, phys_inds, phys=None):
        """positions of the in-annulus grid points inside the compact physical vector.  `ia_is_phys`: a point within
        round-off of the boundary can carry annular coordinates and still be classified exterior by the inside test;
        it has no slot in the compact vector (phys_inds is 0 there) and must not be written."""
        self.ia_inds = phys_inds[self.grid_ia_xind, self.grid_ia_yind]
        self.ia_is_phys = (np.ones(self.ia_inds.shape, dtype=bool) if phys is None
                           else phys[self.grid_ia_xind, self.grid_ia_yind])

    ############################################################################
    def _generate_radial_grid(self):
        bdy = self.bdy
        X, Y = bdy.x, bdy.y
        N = bdy.N
        NX, NY = bdy.normal_x, bdy.normal_y
        T, DT = bdy.t, bdy.dt
        sign = -1 if self.interior else 1
        self.interface = GSB(x=X + sign*self.radial_width*NX, y=Y + sign*self.radial_width*NY)
        lb = -self.radial_width if self.interior else 0.0
        ub = 0.0 if self.interior else self.radial_width
        rc, rv, rat = get_chebyshev_nodes(lb, ub, self.M)
        self.radial_rv = rv
        self.radial_tv = T
        self.radial_r, self.radial_t = np.meshgrid(rv, T, indexing='ij')
        self.radial_x = X + self.radial_r*NX
        self.radial_y = Y + self.radial_r*NY
        OX, OY = self.interface.x, self.interface.y
        self._minx = X.min() if self.interior else OX.min()
        self._maxx = X.max() if self.interior else OX.max()
        self._miny = Y.min() if self.interior else OY.min()
        self._maxy = Y.max() if self.interior else OY.max()
        self.radial_k = np.fft.fftfreq(N, DT/(2*np.pi))
        self.radial_k_filter = SimpleFourierFilter(self.radial_k, 'fraction', fraction=2/3.)
        self.radial_speed = bdy.speed*(1.0 + bdy.curvature*self.radial_r)
        self.inverse_radial_speed = 1.0/self.radial_speed
        self.interpolation_k = np.fft.fftfreq(2*self.M, 1/(2*self.M))
        self.interpolation_modifier = np.exp(-1j*self.interpolation_k[:, None]*np.pi/self.M/2)
        self.radial_shape = (self.M, N)
        V0 = np.polynomial.chebyshev.chebvander(rc, self.M - 1)
        VI0 = np.linalg.inv(V0)
        DC01 = np.polynomial.chebyshev.chebder(np.eye(self.M))/rat
        DC00 = np.vstack([DC01, np.zeros(self.M)])
        self.D00 = V0.dot(DC00.dot(VI0))
        p1 = np.polynomial.chebyshev.chebvander(1, self.M - 1).dot(VI0)[0]
        m1 = np.polynomial.chebyshev.chebvander(-1, self.M - 1).dot(VI0)[0]
        if self.interior:
            self.chebyshev_interp_f_to_bdy = p1
            self.chebyshev_interp_f_to_interface = m1
        else:
            self.chebyshev_interp_f_to_bdy = m1
            self.chebyshev_interp_f_to_interface = p1
        self.chebyshev_interp_df_dn_to_bdy = self.chebyshev_interp_f_to_bdy.dot(self.D00)
        self.chebyshev_interp_df_dn2_to_bdy = self.chebyshev_interp_f_to_bdy.dot(self.D00.dot(self.D00))
        self.bdy_centroid_x = np.mean(bdy.x)
        self.bdy_centroid_y = np.mean(bdy.y)
        self.approximate_radius = np.mean(np.hypot(bdy.x - self.bdy_centroid_x,
                                                   bdy.y - self.bdy_centroid_y))
        self.radial_targ = PointSet(self.radial_x.ravel(), self.radial_y.ravel())
        _, w = fejer_1(self.M)
        self.radial_quadrature_weights = bdy.dt*w[:, None]*self.radial_width/2.0*self.radial_speed

    def check_if_r_in_annulus(self, r):
        if self.interior:
            w1 = r <= 0
            w2 = r >= -self.radial_width
        else:
            w1 = r >= 0
            w2 = r <= self.radial_width
        return np.logical_and(w1, w2), np.logical_not(w2), np.logical_not(w1)

    def nufft_transform_r(self, r):
        lb = -self.radial_width if self.interior else 0.0
        ub = 0.0 if self.interior else self.radial_width
        return np.arccos(np.clip(affine_transformation(r, lb, ub, 1.0, -1.0), -1.0, 1.0))

    def fix_r(self, r):
        if self.interior:
            r[r > 0.0] = 0.0
        else:
            r[r < 0.0] = 0.0

    ############################################################################
    def interpolate_radial_to_boundary(self, f):
        return self.chebyshev_interp_f_to_bdy.dot(f)
    def interpolate_radial_to_interface(self, f):
        return self.chebyshev_interp_f_to_interface.dot(f)
    def interpolate_radial_to_boundary_normal_derivative(self, f):
        return self.chebyshev_interp_df_dn_to_bdy.dot(f)
    def interpolate_radial_to_boundary_normal_derivative2(self, f):
        return self.chebyshev_interp_df_dn2_to_bdy.dot(f)
    def interpolate_radial_to_boundary_any_r_derivative(self, f, d):
        if d == 0:
            return self.interpolate_radial_to_boundary(f)
        DER = np.linalg.matrix_power(self.D00, d)
        return self.interpolate_radial_to_boundary(DER.dot(f))

    ############################################################################
    def interpolate_radial_to_points(self, fr, transf_r, t):
        """
        Chebyshev x Fourier interpolation of the radial function fr to points with
        coordinates (arccos-mapped r, t)  (embedded_boundary.py:419-434).  Runs on
        the device through the radial->points kernel.
        """
        from .device import radial_to_points
        return radial_to_points(np.asarray(fr, dtype=float), np.asarray(transf_r, dtype=float),
                                np.asarray(t, dtype=float))

    def interpolate_radial_to_grid1(self, fr, f):
        vals = self.interpolate_radial_to_points(fr, self.grid_ia_r_transf, self.grid_ia_t)
        f[self.grid_ia_xind, self.grid_ia_yind] = vals

    def interpolate_radial_to_grid2(self, fr, f):
        vals = self.interpolate_radial_to_points(fr, self.grid_ia_r_transf, self.grid_ia_t)
        ok = self.ia_is_phys
        f[self.ia_inds[ok]] = vals[ok]

    ############################################################################
    def _radial_grid_t_derivative(self, f):
        return np.fft.ifft(np.fft.fft(f)*1j*self.radial_k).real
    def _radial_grid_tau_derivative(self, f):
        return self._radial_grid_t_derivative(f)*self.inverse_radial_speed
    def _radial_grid_r_derivative(self, f):
        return self.D00.dot(f)
    def radial_grid_derivatives(self, f):
        bdy = self.bdy
        ft = self._radial_grid_tau_derivative(f)
        fr = self._radial_grid_r_derivative(f)
        return fr*bdy.normal_x + ft*bdy.tangent_x, fr*bdy.normal_y + ft*bdy.tangent_y
    def radial_grid_laplacian(self, f):
        fx, fy = self.radial_grid_derivatives(f)
        fxx, _ = self.radial_grid_derivatives(fx)
        _, fyy = self.radial_grid_derivatives(fy)
        return fxx + fyy
    def gradient(self, fx, fy, fr):
        fxr, fyr = self.radial_grid_derivatives(fr)
        self.interpolate_radial_to_grid1(fxr, fx)
        self.interpolate_radial_to_grid1(fyr, fy)
        return fxr, fyr
    def laplacian(self, lapf, fr):
        lapfr = self.radial_grid_laplacian(fr)
        self.interpolate_radial_to_grid1(lapfr, lapf)
        return lapfr

    ############################################################################
    def convert_uv_to_rt(self, fu, fv):
        bdy = self.bdy
        return fu*bdy.normal_x + fv*bdy.normal_y, fu*bdy.tangent_x + fv*bdy.tangent_y
    def convert_rt_to_uv(self, fr, ft):
        bdy = self.bdy
        return fr*bdy.normal_x + ft*bdy.tangent_x, fr*bdy.normal_y + ft*bdy.tangent_y

    ############################################################################
    def radial_integral(self, fr):
        return np.sum(fr*self.radial_cutoff[:, None]*self.radial_quadrature_weights)
