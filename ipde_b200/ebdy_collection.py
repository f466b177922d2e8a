"""
EmbeddedBoundaryCollection: several EmbeddedBoundary objects over one periodic
background grid.

API mirror of ipde/ebdy_collection.py:230-829: `register_grid` (:351-526),
`generate_grid` (:279-339), `ready_bump` / `demean_function` (:796-810),
`divide_pnai` / `v2l` / `v2r` (:541-563), `interpolate_grid_to_interface`
(:578-614), `interpolate_radial_to_grid1` (:619-620), `grid_integral` /
`volume_integral` (:814-829) and the attribute set the solvers and example
scripts read (phys, ext, grid_step, in_annulus, phys_not_in_annulus, grid_phys,
grid_pna, grid_pnai, radial_pts, grid_and_radial_pts, all_bvx/y, all_ivx/y,
interfaces_x/y_transf, kx, ky, ikx, iky, lap, splitter, rsplitter, bumpy ...).

Registration is setup work and stays in NumPy.  The per-solve members
(interpolate_grid_to_interface, interpolate_radial_to_grid1, demean_function)
are executed by the CUDA library (ipde_b200/device.py).
"""
import numpy as np
from .geometry import PointSet, LazyPointSet, Grid
from .near import points_inside_curve, CurveTaylor, compute_local_coordinates
from .embedded_boundary import EmbeddedBoundary, LoadEmbeddedBoundary
from .embedded_function import EmbeddedFunction, BoundaryFunction
from .utilities import affine_transformation


class EmbeddedPointPartition(object):
    """
    Unstructured points (x, y) sorted by where an EmbeddedFunction is to be read there
    (API mirror of ipde/ebdy_collection.py:37-229):

        zone 1   not in any annulus and not flagged exterior   -> trigonometric interpolation of the smoothed grid part
        zone 2   inside the annulus of boundary i (zone2l[i])  -> Chebyshev x Fourier interpolation of radial block i,
                 with the coordinates (zone2r, zone2_transfr, zone2t)
        zone 3   within a radial width of boundary i on its non-physical side (zone3l[i])  -> nan

    As in the reference, only points within one radial width of a boundary get coordinates; a point far from every
    boundary is taken to be physical (zone 1).  Coordinates: closest node of the 2x upsampled curve from a k-d tree
    as the guess, then the Newton iteration of near.py on the Taylor tables of the trigonometric interpolant --
    the machinery grid registration uses (replaces near_finder.points_near_curve_coarse + compute_local_coordinates,
    :109-121).  dzl / gil (danger-zone and guess-index lists, :59-66) are honoured when given.
    """
    def __init__(self, ebdyc, x, y, fix_r=False, dzl=None, gil=None):
        from scipy.spatial import cKDTree
        self.grid = ebdyc.grid
        self.x_check, self.y_check, self.fix_r = x, y, fix_r
        self.dzl, self.gil = dzl, gil
        self.sh = np.shape(x)
        self.x = np.asarray(x, dtype=float).ravel()
        self.y = np.asarray(y, dtype=float).ravel()
        self.size = n = self.x.size
        z = self.x + 1j*self.y
        zone1 = np.ones(n, dtype=bool)
        zone1_or_2 = np.ones(n, dtype=bool)
        self.zone2l, self.zone2r, self.zone2_transfr, self.zone2t = [], [], [], []
        self.zone3l, self.zone3r, self.zone3t = [], [], []
        self.full_t = np.full(n, np.nan)
        self.full_r = np.full(n, np.nan)
        for ind, ebdy in enumerate(ebdyc):
            ct = CurveTaylor(ebdy.bdy.c)
            if dzl is not None and gil is not None and dzl[ind] is not None and gil[ind] is not None:
                near = np.asarray(dzl[ind], dtype=np.int64)
                guess = 2*np.asarray(gil[ind], dtype=np.int64)          # boundary node -> node of the 2x curve
            else:
                nodes = ct.tab[0]
                tree = cKDTree(np.column_stack([nodes.real, nodes.imag]))
                # a point within `width` of the curve is within sqrt(width^2 + (1 + kappa width) (hmax/2)^2) of a node
                hmax = np.abs(np.diff(np.append(nodes, nodes[0]))).max()
                dist, guess = tree.query(np.column_stack([self.x, self.y]), workers=-1)
                near = np.flatnonzero(dist <= np.hypot(ebdy.radial_width, 0.75*hmax))
                guess = guess[near]
            if near.size == 0:
                e = np.zeros(0)
                ei = np.zeros(0, dtype=np.int64)
                self.zone2l.append(ei); self.zone2r.append(e); self.zone2_transfr.append(e); self.zone2t.append(e)
                self.zone3l.append(ei); self.zone3r.append(e); self.zone3t.append(e)
                continue
            t, r = compute_local_coordinates(ct, z[near], guess, tol=getattr(ebdy, 'coordinate_tolerance', 1e-14))
            if fix_r:
                ebdy.fix_r(r)
            self.full_t[near], self.full_r[near] = t, r
            inside, phys_beyond, exterior = ebdy.check_if_r_in_annulus(r)
            zone1[near] &= phys_beyond
            zone1_or_2[near] &= ~exterior
            self.zone2l.append(near[inside])
            self.zone2r.append(r[inside])
            self.zone2_transfr.append(ebdy.nufft_transform_r(r[inside]))
            self.zone2t.append(t[inside])
            self.zone3l.append(near[exterior])
            self.zone3r.append(r[exterior])
            self.zone3t.append(t[exterior])
        self.zone1, self.zone1_or_2 = zone1, zone1_or_2
        self.zone1_N = int(zone1.sum())
        self.zone2_Ns = [len(a) for a in self.zone2l]
        self.zone3_Ns = [len(a) for a in self.zone3l]
        self.zone2_N, self.zone3_N = int(np.sum(self.zone2_Ns)), int(np.sum(self.zone3_Ns))
        g = self.grid
        self.x_transf = affine_transformation(self.x[zone1], g.x_bounds[0], g.x_bounds[1], 0.0, 2*np.pi)
        self.y_transf = affine_transformation(self.y[zone1], g.y_bounds[0], g.y_bounds[1], 0.0, 2*np.pi)

    def is_it_i(self, x, y, fix_r):
        return x is self.x_check and y is self.y_check and fix_r == self.fix_r

    def reshape(self, out):
        return out.reshape(self.sh)

    def get_Ns(self):
        return self.zone1_N, self.zone2_N, self.zone3_N


def LoadEmbeddedBoundaryCollection(d):
    ebdy_list = [LoadEmbeddedBoundary(e) for e in d['ebdy_list']]
    ebdyc = EmbeddedBoundaryCollection(ebdy_list)
    if d['grid'] is not None:
        ebdyc.register_grid(Grid(**d['grid']), danger_zone_distance=d['ddd'])
    if d['bumpy'] is not None:
        ebdyc.bumpy = d['bumpy']
        ebdyc.bumpy_readied = True
    return ebdyc


class EmbeddedBoundaryCollection(object):
    def __init__(self, ebdy_list):
        self.ebdys = ebdy_list
        self.N = len(self.ebdys)
        self.registered_partitions = []
        self.bumpy_readied = False
        self.bump_location = None
        self._device = None

    def __iter__(self):
        return (ebdy for ebdy in self.ebdys)
    def __len__(self):
        return len(self.ebdys)
    def __getitem__(self, ind):
        return self.ebdys[ind]

    def save(self):
        ebdy_list = [ebdy.save() for ebdy in self.ebdys]
        if hasattr(self, 'grid'):
            g = self.grid
            grid = {'x_bounds': g.x_bounds, 'y_bounds': g.y_bounds, 'Nx': g.Nx, 'Ny': g.Ny,
                    'mask': g.mask, 'x_endpoints': g.x_endpoints, 'y_endpoints': g.y_endpoints}
            ddd = self.danger_zone_distance
        else:
            grid, ddd = None, None
        return {'ebdy_list': ebdy_list, 'grid': grid,
                'bumpy': self.bumpy if hasattr(self, 'bumpy') else None, 'ddd': ddd}

    def generate_grid(self, h=None, Ns=None, force_square=False, danger_zone_distance=None):
        """Auto-generate the background grid around the first (interior) boundary (:279-339)."""
        iebdy = self[0]
        if not iebdy.interior:
            raise Exception('Generate grid may only be used if the first boundary is interior.')
        ibdy = iebdy.bdy
        cheat_space = iebdy.radial_width
        xmin = ibdy.x.min() - cheat_space
        ymin = ibdy.y.min() - cheat_space
        xmax = ibdy.x.max() + 2*cheat_space
        ymax = ibdy.y.max() + 2*cheat_space
        self.bump_location = [ibdy.x.max() + cheat_space, ibdy.y.max() + cheat_space]
        xran, yran = xmax - xmin, ymax - ymin
        if h is None:
            h = iebdy.radial_width/iebdy.M
        if Ns is None:
            Nx = 2*int(0.5*np.ceil(xran/h))
            Ny = 2*int(0.5*np.ceil(yran/h))
        else:
            Nx, Ny = Ns
            if xmin + Nx*h < xmin + xran:
                raise Exception('Provided value of Nx is too small')
            if ymin + Ny*h < ymin + yran:
                raise Exception('Provided value of Ny is too small')
        if force_square:
            Nx = Ny = max(Nx, Ny)
        if Ns is None:
            # the device FFT handles lengths with prime factors <= 13: an automatically chosen size outside that set
            # is padded on the high side to the next friendly one (same origin, same spacing; only more exterior cells)
            from ._lib import lib
            ok = lambda n: lib().ipde_fft_size_supported(n) and lib().ipde_fft_size_supported(n//2)
            if not ok(Nx):
                Nx = 2*lib().ipde_fft_next_size(-(-Nx//2))
            if not ok(Ny):
                Ny = 2*lib().ipde_fft_next_size(-(-Ny//2))
            if force_square:
                Nx = Ny = max(Nx, Ny)
        grid = Grid([xmin, xmin + Nx*h], Nx, [ymin, ymin + Ny*h], Ny,
                    x_endpoints=[True, False], y_endpoints=[True, False])
        assert np.abs(grid.xh - h) < 1e-15, 'Gridspacing not what was requested'
        assert np.abs(grid.yh - h) < 1e-15, 'Gridspacing not what was requested'
        self.register_grid(grid, danger_zone_distance=danger_zone_distance)
        self.bumpy_readied = False
        return grid

    def register_grid(self, grid, danger_zone_distance=None, verbose=False):
        self.grid = grid
        self.danger_zone_distance = danger_zone_distance
        self._device = None
        for ei, ebdy in enumerate(self):
            if verbose:
                print('Registering ebdy #', ei + 1, 'of', self.N)
            ebdy.register_grid(grid, index=ei + 1,
                               danger_zone_distance=danger_zone_distance, verbose=verbose)
            ebdy.ebdyc = self          # lets the single-boundary adapters find the collection an ebdy belongs to
        # physical region: inside the interior boundary, outside every exterior one
        phys = np.ones(grid.shape, dtype=bool)
        for ebdy in self:
            inside = points_inside_curve(ebdy.bdy, grid.xv, grid.yv, ebdy.near_curve_result)
            phys = np.logical_and(phys, inside if ebdy.interior else np.logical_not(inside))
        self.phys = phys
        self.ext = np.logical_not(phys)
        self.phys_inds = np.zeros(grid.shape, dtype=int)
        self.phys_n = int(np.sum(phys))
        self.phys_inds[phys] = np.arange(self.phys_n)

        kxv = np.fft.fftfreq(grid.Nx, grid.xh/(2*np.pi))
        kyv = np.fft.fftfreq(grid.Ny, grid.yh/(2*np.pi))
        self.kx = kxv[:, None]
        self.ky = kyv
        self.ikx = 1j*self.kx
        self.iky = 1j*self.ky
        self.lap = -self.kx*self.kx - self.ky*self.ky

        iax = np.concatenate([ebdy.grid_ia_xind for ebdy in self])
        iay = np.concatenate([ebdy.grid_ia_yind for ebdy in self])
        ia = np.zeros(grid.shape, dtype=bool)
        ia[iax, iay] = True
        self.in_annulus = ia
        self.phys_not_in_annulus = np.logical_and(phys, np.logical_not(ia))
        for ebdy in self:
            ebdy.register_ia_inds(self.phys_inds, self.phys)

        self.grid_step = phys.astype(float)
        for ebdy in self:
            self.grid_step[ebdy.grid_ia_xind, ebdy.grid_ia_yind] *= ebdy.grid_to_radial_step

        # Collection-wide point sets (ipde/ebdy_collection.py:414-446).  Each is O(N^2) doubles; the device path needs
        # only their sizes, so the coordinates are produced on first access (LazyPointSet) from the index lists.
        xv, yv = grid.xv, grid.yv
        def masked(mask):
            def build():
                ix, iy = np.nonzero(mask)
                return xv[ix], yv[iy]
            return build
        self.radial_x = np.concatenate([ebdy.radial_x.ravel() for ebdy in self])
        self.radial_y = np.concatenate([ebdy.radial_y.ravel() for ebdy in self])
        self.radial_pts = PointSet(self.radial_x, self.radial_y)
        self.all_bvx = np.concatenate([ebdy.bdy.x for ebdy in self])
        self.all_bvy = np.concatenate([ebdy.bdy.y for ebdy in self])
        self.all_bv = PointSet(self.all_bvx, self.all_bvy)
        self.all_ivx = np.concatenate([ebdy.interface.x for ebdy in self])
        self.all_ivy = np.concatenate([ebdy.interface.y for ebdy in self])
        self.all_iv = PointSet(self.all_ivx, self.all_ivy)
        n_phys = self.phys_n
        n_pna = int(np.sum(self.phys_not_in_annulus))
        n_ext = int(phys.size - n_phys)
        nr, ni = self.radial_x.size, self.all_ivx.size
        def plus(build, ex, ey):
            def b():
                x, y = build()
                return np.concatenate([x, ex]), np.concatenate([y, ey])
            return b
        self.grid_phys = LazyPointSet(n_phys, masked(phys))
        self.grid_and_radial_pts = LazyPointSet(n_phys + nr, plus(masked(phys), self.radial_x, self.radial_y))
        self.grid_pna = LazyPointSet(n_pna, masked(self.phys_not_in_annulus))
        self.grid_pna_num = n_pna
        self.pnar = LazyPointSet(n_pna + nr, plus(masked(self.phys_not_in_annulus), self.radial_x, self.radial_y))
        self.exterior_targ = LazyPointSet(n_ext, masked(self.ext))
        self.grid_pnai = LazyPointSet(n_pna + ni, plus(masked(self.phys_not_in_annulus), self.all_ivx, self.all_ivy))

        self.bdy_Ns = [ebdy.bdy.N for ebdy in self]
        self.cumsum_bdy_Ns = np.cumsum(self.bdy_Ns)
        self.splitter = self.cumsum_bdy_Ns[:-1]
        self.radial_Ns = [ebdy.bdy.N*ebdy.M for ebdy in self]
        self.cumsum_radial_Ns = np.cumsum(self.radial_Ns)
        self.rsplitter = self.cumsum_radial_Ns[:-1]
        self.interfaces_x_transf = np.concatenate([ebdy.interface_x_transf for ebdy in self])
        self.interfaces_y_transf = np.concatenate([ebdy.interface_y_transf for ebdy in self])
        self.grid_dof = self.grid_phys.N
        self.radial_dof_list = [int(np.prod(ebdy.radial_shape)) for ebdy in self]
        self.radial_dof = int(np.sum(self.radial_dof_list))
        self.dof = self.grid_dof + self.radial_dof

    @property
    def exterior_x(self):
        return self.exterior_targ.x

    @property
    def exterior_y(self):
        return self.exterior_targ.y

    ############################################################################
    def divide_grid_and_radial(self, v):
        g, w = np.split(v, [self.grid_phys.N])
        return g, self.v2r(w)
    def divide_pnai(self, v):
        g, w = np.split(v, [self.grid_pna.N])
        return g, self.v2l(w)
    def v2l(self, v):
        return np.split(v, self.splitter)
    def v2l2(self, v):
        return np.split(v, self.splitter*2)
    def v2l_vector(self, v):
        return np.split(v, self.splitter, axis=1)
    def v2r(self, v):
        return [L.reshape(ebdy.radial_x.shape) for L, ebdy in zip(np.split(v, self.rsplitter), self.ebdys)]

    ############################################################################
    @property
    def device(self):
        """Device-resident copy of the registered geometry (built on first use)."""
        if self._device is None:
            from .device import DeviceCollection
            self._device = DeviceCollection(self)
        return self._device

    def interpolate_radial_to_boundary(self, f):
        out = BoundaryFunction(self)
        out.load_data([ebdy.interpolate_radial_to_boundary(fr) for ebdy, fr in zip(self.ebdys, f)])
        return out

    def interpolate_grid_to_interface(self, f, order=3, cutoff=None):
        """
        Interpolate a grid function to all interface points (:578-614): order 1..7 =
        periodic piecewise polynomial, np.inf = trigonometric (type-2 NUFFT).  Runs
        on the device.
        """
        if type(f) == EmbeddedFunction:
            f = f.get_grid_value()
        if cutoff is None:
            cutoff = order == np.inf
        fc = f*self.grid_step if cutoff else f
        return self.device.interpolate_grid_to_interface(fc, order)

    def interpolate_radial_to_grid1(self, fr_list, f=None):
        return [ebdy.interpolate_radial_to_grid1(fr, f) for ebdy, fr in zip(self.ebdys, fr_list)]
    def update_radial_to_grid1(self, f):
        _ = self.interpolate_radial_to_grid1(f.get_radial_value_list(), f['grid'])

    ############################################################################
    # EmbeddedFunction -> unstructured points (ipde/ebdy_collection.py:651-707)
    def get_interpolation_key(self, x, y, fix_r=False, dzl=None, gil=None):
        for key, part in enumerate(getattr(self, 'registered_partitions', [])):
            if part.is_it_i(x, y, fix_r):
                return key
        return self.register_points(x, y, fix_r, dzl, gil)

    def register_points(self, x, y, fix_r=False, dzl=None, gil=None):
        if not hasattr(self, 'registered_partitions'):
            self.registered_partitions = []
        self.registered_partitions.append(EmbeddedPointPartition(self, x, y, fix_r, dzl, gil))
        return len(self.registered_partitions) - 1

    def interpolate_to_points(self, ff, x, y, fix_r=False, dzl=None, gil=None):
        """
        Values of the EmbeddedFunction ff at the points (x, y): type-2 NUFFT of the smoothed grid part in zone 1 (the
        reference's finufft.nufft2d2 call, :687-694), Chebyshev x Fourier interpolation of the radial blocks in the
        annuli (:695-702), nan on the flagged exterior points (:704-706).  Device kernels: scatter + cutoff, the grid
        FFT, k_nufft_pad / k_nufft_interp, k_radial_interp.
        """
        from .device import to_dev, ptr, stream, Nufft2
        from ._lib import lib
        p = self.registered_partitions[self.get_interpolation_key(x, y, fix_r, dzl, gil)]
        out = np.empty(p.size)
        if p.zone1_N > 0:
            D = self.device
            f_dev = to_dev(np.ascontiguousarray(np.asarray(ff, dtype=float).reshape(-1)))
            lib().ipde_grid_scatter_cutoff(D.grid.h, ptr(f_dev), ptr(D.phys_lin), ptr(D.step_compact), D.n_phys,
                                           ptr(D.fc), stream())
            spec = D.grid.fft2(D.fc)
            if getattr(p, '_nufft', None) is None or p._nufft.grid is not D.grid:
                p._nufft = Nufft2(D.grid, p.x_transf, p.y_transf)
            out[p.zone1] = p._nufft(spec).cpu().numpy()
        for ind, ebdy in enumerate(self.ebdys):
            if p.zone2_Ns[ind] > 0:
                out[p.zone2l[ind]] = ebdy.interpolate_radial_to_points(ff[ind], p.zone2_transfr[ind], p.zone2t[ind])
        for z3 in p.zone3l:
            out[z3] = np.nan
        return out

    ############################################################################
    # calculus on EmbeddedFunctions (ipde/ebdy_collection.py:708-792), on the device: ipde_b200/calculus.py
    def _calculus(self):
        if getattr(self, '_calc', None) is None or self._calc.D is not self.device:
            from .calculus import DeviceCalculus
            self._calc = DeviceCalculus(self)
        return self._calc

    def gradient(self, ff, derivative_type='spectral'):
        """(fx, fy) EmbeddedFunctions; 'spectral' (cutoff + Fourier on the grid) or 'fourth' (centred differences)"""
        from .device import to_dev
        gx, gy = self._calculus().gradient_device(to_dev(np.asarray(ff, dtype=float).reshape(-1)), derivative_type)
        return (EmbeddedFunction(self, linear_data=gx.cpu().numpy()), EmbeddedFunction(self, linear_data=gy.cpu().numpy()))

    def laplacian(self, ff, derivative_type='spectral'):
        from .device import to_dev
        lap = self._calculus().laplacian_device(to_dev(np.asarray(ff, dtype=float).reshape(-1)), derivative_type)
        return EmbeddedFunction(self, linear_data=lap.cpu().numpy())

    ############################################################################
    def ready_bump(self, bump, bump_loc=None, bump_width=None):
        if bump_width is None:
            bump_width = self[0].radial_width
        if bump_loc is None:
            if self.bump_location is None:
                raise Exception('if ebdyc has no bump_location, need to give bump_loc')
            bump_loc = self.bump_location
        grr = np.hypot(self.grid.xg - bump_loc[0], self.grid.yg - bump_loc[1])
        bumpy = bump(affine_transformation(grr, 0, bump_width, 0, 1))
        self.bumpy = bumpy/self.grid_integral(bumpy)
        self.bumpy_readied = True
        self._device = None

    def demean_function(self, f):
        return f - self.grid_integral(f)*self.bumpy

    def grid_integral(self, f):
        return np.sum(f)*self.grid.xh*self.grid.yh

    def volume_integral(self, f):
        if f.ebdyc() is not self:
            raise Exception('EmbeddedFunction must be defined over this EmbeddedBoundaryCollection.')
        integral = self.grid_integral(f.get_smoothed_grid_value())
        for fr, ebdy in zip(f, self):
            integral += ebdy.radial_integral(fr)
        return integral
