"""
Host-side (NumPy) geometry containers used at *setup* time.

The reference takes these from the un-vendored `pybie2d` package (SURVEY.md
section 2.2): `Global_Smooth_Boundary` (reference call sites
ipde/embedded_boundary.py:9,290,354), `Grid` (ipde/ebdy_collection.py:20,329),
`PointSet` (ipde/embedded_boundary.py:10), `BoundaryCollection`
(ipde/solvers/multi_boundary/scalar.py:4,35-39) and the curve generators
`star` / `squished_circle` (examples/multi_modified_helmholtz.py:34-36).
pybie2d is not in the container, so this is a restatement of the documented
behaviour: a closed curve sampled at N equispaced parameter values, spectral
differentiation for tangents/normals/curvature, trapezoid weights.

Sign conventions pinned by the reference's use (embedded_boundary.py:299-319):
`normal` is the outward normal of a counter-clockwise curve and curvature is
positive for a CCW circle, so that `speed*(1+curvature*r)` is the speed of the
curve offset a distance r along the normal.

Nothing here is on the per-solve path.
"""
import numpy as np


def star(N, x=0.0, y=0.0, r=1.0, a=0.5, f=3, rot=0.0):
    """c(t) = (x+iy) + r(1 + a cos(f(t-rot))) e^{it}"""
    t = np.linspace(0.0, 2.0*np.pi, N, endpoint=False)
    return (x + 1j*y) + (r + r*a*np.cos(f*(t - rot)))*np.exp(1j*t)


def squished_circle(N, x=0.0, y=0.0, r=1.0, b=1.0, rot=0.0):
    """ellipse with semi-axes r and r*b rotated by rot"""
    t = np.linspace(0.0, 2.0*np.pi, N, endpoint=False)
    return (x + 1j*y) + np.exp(1j*rot)*r*(np.cos(t) + 1j*b*np.sin(t))


class PointSet(object):
    """Bare collection of points (x, y); `c` = x + iy (built on first use)."""
    def __init__(self, x=None, y=None, c=None):
        if c is not None:
            c = np.asarray(c)
            x, y = c.real.copy(), c.imag.copy()
        self.x = np.ascontiguousarray(x, dtype=float).ravel()
        self.y = np.ascontiguousarray(y, dtype=float).ravel()
        self.N = self.x.size

    @property
    def c(self):
        return self.x + 1j*self.y

    def get_stacked_boundary(self, T=True):
        return np.vstack([self.x, self.y])


class LazyPointSet(object):
    """
    PointSet whose coordinates are produced on first access.  The collection-wide point sets of a registered grid
    (grid_phys, grid_pna, grid_pnai, exterior_targ, ... -- ipde/ebdy_collection.py:414-446) are O(N^2) each; at
    16384^2 they add up to ~25 GB of host memory, and the device path needs only their sizes.
    """
    def __init__(self, n, build):
        self.N = int(n)
        self._build = build
        self._xy = None

    def _get(self):
        if self._xy is None:
            x, y = self._build()
            self._xy = (np.ascontiguousarray(x, dtype=float).ravel(), np.ascontiguousarray(y, dtype=float).ravel())
            assert self._xy[0].size == self.N
        return self._xy

    @property
    def x(self):
        return self._get()[0]

    @property
    def y(self):
        return self._get()[1]

    @property
    def c(self):
        return self.x + 1j*self.y

    def get_stacked_boundary(self, T=True):
        return np.vstack([self.x, self.y])


class GlobalSmoothBoundary(PointSet):
    """
    Smooth closed curve given at N equispaced parameter values t_j = 2 pi j / N.
    Attribute names follow pybie2d's Global_Smooth_Boundary as used by the
    reference: x, y, c, N, t, dt, k, normal_x/y, tangent_x/y, speed, curvature,
    weights, cp, cpp.
    """
    def __init__(self, x=None, y=None, c=None):
        super().__init__(x, y, c)
        N = self.N
        self.t, self.dt = np.linspace(0.0, 2.0*np.pi, N, endpoint=False, retstep=True)
        self.k = np.fft.fftfreq(N, 1.0/N)
        self.ik = 1j*self.k
        self.chat = np.fft.fft(self.c)
        self.cp = np.fft.ifft(self.chat*self.ik)
        self.cpp = np.fft.ifft(self.chat*self.ik*self.ik)
        self.speed = np.abs(self.cp)
        self.tangent_c = self.cp/self.speed
        self.tangent_x = self.tangent_c.real.copy()
        self.tangent_y = self.tangent_c.imag.copy()
        self.normal_c = -1j*self.tangent_c
        self.normal_x = self.normal_c.real.copy()
        self.normal_y = self.normal_c.imag.copy()
        self.curvature = -(np.conj(self.cpp)*self.normal_c).real/self.speed**2
        self.weights = self.dt*self.speed
        self.complex_weights = self.dt*self.cp
        self.max_h = np.max(self.weights)
        self.area = 0.5*np.sum((self.x*self.normal_x + self.y*self.normal_y)*self.weights)
        self.perimeter = np.sum(self.weights)

    def fourier_resample(self, n):
        """Trigonometric interpolation of the curve to n equispaced points."""
        return fourier_resample(self.c, n)


GSB = GlobalSmoothBoundary
Global_Smooth_Boundary = GlobalSmoothBoundary


def fourier_resample(f, n):
    """Resample periodic samples f (length N, real or complex) to n samples."""
    f = np.asarray(f)
    N = f.shape[0]
    if n == N:
        return f.copy()
    fh = np.fft.fft(f, axis=0)
    out = np.zeros((n,) + f.shape[1:], dtype=complex)
    m = min(N, n)
    h = (m - 1)//2          # modes 1..h and -1..-h are copied verbatim
    out[:h+1] = fh[:h+1]
    if h > 0:
        out[n-h:] = fh[N-h:]
    if m % 2 == 0:
        q = m//2
        if n > N:            # upsampling: split the Nyquist mode of the coarse grid
            out[q] = 0.5*fh[q]
            out[n-q] = 0.5*fh[q]
        else:                # downsampling: fold +-q onto the Nyquist mode
            out[q] = fh[q] + (fh[N-q] if N > n else 0.0)
    res = np.fft.ifft(out, axis=0)*(float(n)/N)
    if np.isrealobj(f):
        res = res.real
    return res


def fourier_resample_matrix(N, n):
    """(n x N) real matrix performing `fourier_resample` on length-N vectors."""
    return fourier_resample(np.eye(N), n)


def arc_length_parameterize(x, y, tol=1e-14, maxiter=50):
    """
    Resample a closed curve at equal arc-length increments (personal_utilities.arc_length_reparametrization.
    arc_length_parameterize as the reference's examples use it, e.g. examples/interior_stokes_with_multi.py:54-55:
    `bdy = GSB(*arc_length_parameterize(bdy.x, bdy.y))`).  Spectral: arc length from the Fourier series of the speed,
    Newton iteration for the parameters, trigonometric interpolation of the curve.  Returns (x, y).
    """
    c = np.asarray(x, dtype=float) + 1j*np.asarray(y, dtype=float)
    N = c.size
    k = np.fft.fftfreq(N, 1.0/N)
    ch = np.fft.fft(c)/N
    sp = np.abs(np.fft.ifft(np.fft.fft(c)*1j*k))
    sh = np.fft.fft(sp)/N
    nz = k != 0
    def L(t):
        E = np.exp(1j*np.outer(t, k[nz]))
        return (sh[0]*t + ((E - 1.0)*(sh[nz]/(1j*k[nz]))).sum(axis=1)).real
    def S(t):
        return (np.exp(1j*np.outer(t, k)).dot(sh)).real
    total = 2*np.pi*sh[0].real
    target = total*np.arange(N)/N
    t = 2*np.pi*np.arange(N)/N
    for _ in range(maxiter):
        dt = (L(t) - target)/S(t)
        t = t - dt
        if np.abs(dt).max() < tol:
            break
    cn = np.exp(1j*np.outer(t, k)).dot(ch)
    return cn.real.copy(), cn.imag.copy()


class Grid(object):
    """
    Uniform tensor grid, following pybie2d.grid.Grid as the reference uses it
    (ebdy_collection.py:329: Grid([xmin,xmax], Nx, [ymin,ymax], Ny,
    x_endpoints=[True,False], y_endpoints=[True,False])).  Arrays are (Nx, Ny)
    with indexing='ij' (x index slow).
    """
    def __init__(self, x_bounds, Nx, y_bounds, Ny, mask=None,
                 x_endpoints=[True, False], y_endpoints=[True, False]):
        self.x_bounds = list(x_bounds)
        self.y_bounds = list(y_bounds)
        self.Nx, self.Ny = int(Nx), int(Ny)
        self.mask = mask
        self.x_endpoints = list(x_endpoints)
        self.y_endpoints = list(y_endpoints)
        self.xv, self.xh = self._axis(self.x_bounds, self.Nx, self.x_endpoints)
        self.yv, self.yh = self._axis(self.y_bounds, self.Ny, self.y_endpoints)
        self.shape = (self.Nx, self.Ny)
        self._xg = None
        self._yg = None

    @staticmethod
    def _axis(bounds, n, endpoints):
        lo, hi = bounds
        if endpoints[0] and endpoints[1]:
            v, h = np.linspace(lo, hi, n, endpoint=True, retstep=True)
        elif endpoints[0]:
            v, h = np.linspace(lo, hi, n, endpoint=False, retstep=True)
        elif endpoints[1]:
            v, h = np.linspace(lo, hi, n, endpoint=False, retstep=True)
            v = v + h
        else:
            h = (hi - lo)/(n + 1)
            v = lo + h*np.arange(1, n + 1)
        return v, h

    # the meshgrids are large at big N: build lazily
    @property
    def xg(self):
        if self._xg is None:
            self._xg, self._yg = np.meshgrid(self.xv, self.yv, indexing='ij')
        return self._xg

    @property
    def yg(self):
        if self._yg is None:
            self._xg, self._yg = np.meshgrid(self.xv, self.yv, indexing='ij')
        return self._yg


class BoundaryCollection(object):
    """
    Concatenation of several source curves (pybie2d BoundaryCollection as used
    at multi_boundary/scalar.py:35-39).
    """
    def __init__(self):
        self.boundaries = []
        self.sides = []

    def add(self, bdy, side='i'):
        self.boundaries.append(bdy)
        self.sides.append(side)

    def amass_information(self):
        self.x = np.concatenate([b.x for b in self.boundaries])
        self.y = np.concatenate([b.y for b in self.boundaries])
        self.c = self.x + 1j*self.y
        self.weights = np.concatenate([b.weights for b in self.boundaries])
        self.normal_x = np.concatenate([b.normal_x for b in self.boundaries])
        self.normal_y = np.concatenate([b.normal_y for b in self.boundaries])
        self.N = self.x.size
        self.Ns = [b.N for b in self.boundaries]

    def get_stacked_boundary(self, T=True):
        return np.vstack([self.x, self.y])
