"""
Device plumbing: torch CUDA tensors as buffers, thin Python objects around the library's
opaque plans, and the device-resident copy of a registered EmbeddedBoundaryCollection.

PyTorch is used for allocation, host<->device copies and streams only; every computation
goes through libipde_b200.so (ipde_b200/_lib.py).  No CPU fallback exists: constructing any
of these without a CUDA device raises.
"""
import numpy as np
import scipy.special
import torch

from ._lib import lib, IpdeError

F64 = torch.float64
C128 = torch.complex128


def require_cuda():
    if not torch.cuda.is_available():
        raise IpdeError('ipde_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def to_dev(a, dtype=None):
    dev = require_cuda()
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a)
    if dtype is not None:
        t = t.to(dtype)
    return t.to(dev)


def empty(shape, dtype=F64):
    return torch.empty(shape, dtype=dtype, device=require_cuda())


def zeros(shape, dtype=F64):
    return torch.zeros(shape, dtype=dtype, device=require_cuda())


def ptr(t):
    return 0 if t is None else t.data_ptr()


def hptr(a):
    """pointer of a host numpy array (kept alive by the caller)"""
    return a.ctypes.data


class HostStaging(object):
    """
    Host <-> device transfer of the flat EmbeddedFunction vectors for the public entry points.  EmbeddedFunctions
    created by this package live in page-locked memory (embedded_function._allocate) and are copied straight from / to
    the user's array; any other host array goes through a lazily allocated pinned staging buffer.
    """

    def __init__(self):
        self._buf = {}

    def _stage(self, key, n):
        b = self._buf.get(key)
        if b is None or b.numel() != n:
            b = self._buf[key] = torch.empty(n, dtype=F64).pin_memory()
        return b

    def upload(self, key, src, dst):
        t = torch.from_numpy(np.ascontiguousarray(src, dtype=float).reshape(-1))
        if not t.is_pinned():
            hb = self._stage(('in', key), t.numel())
            hb.copy_(t)
            t = hb
        dst.copy_(t, non_blocking=True)

    def download(self, pairs):
        """pairs: [(device vector, host ndarray)]; one synchronisation for all of them"""
        late = []
        for key, (src, out) in enumerate(pairs):
            t = torch.from_numpy(np.asarray(out).reshape(-1))
            if t.is_pinned():
                t.copy_(src, non_blocking=True)
            else:
                hb = self._stage(('out', key), t.numel())
                hb.copy_(src, non_blocking=True)
                late.append((t, hb))
        torch.cuda.current_stream().synchronize()
        for t, hb in late:
            t.copy_(hb)


def stream():
    return torch.cuda.current_stream().cuda_stream


# ----------------------------------------------------------------------------------------
# "exponential of semicircle" kernel (Barnett, Magland, af Klinteberg 2019) for the type-2 NUFFT
def es_kernel_width(eps=1e-14):
    return 16 if eps <= 1e-13 else max(4, int(np.ceil(-np.log10(eps/10.0))) + 1)


def es_deconvolution(w, beta, n, fn):
    """
    factors dec[k], k in FFT order over n coarse modes, such that
    sum_l b_l phi(x - x_l) reproduces exp(i k x) when b = ifft(dec * delta_k) on the fn-point fine grid
    """
    z, wq = np.polynomial.legendre.leggauss(256)
    phi = np.exp(beta*(np.sqrt(1.0 - z*z) - 1.0))
    s = np.abs(np.fft.fftfreq(n, 1.0/n))
    xi = np.pi*w*s/fn
    I = (np.cos(np.outer(xi, z))*phi).dot(wq)
    return 2.0/(w*I)


_k0_table_set = set()


def ensure_k0_table():
    """Chebyshev fits of sqrt(x) e^x K0(x) and sqrt(x) e^x K1(x) in t = 4/x - 1 on x >= 2, uploaded once per device
    (constant memory, csrc/bessel.cuh)."""
    dev = torch.cuda.current_device()
    if dev in _k0_table_set:
        return
    nc = 36
    t, _ = np.polynomial.chebyshev.chebgauss(4*nc)
    x = 4.0/(t + 1.0)
    c0 = np.polynomial.chebyshev.chebfit(t, np.sqrt(x)*scipy.special.k0e(x), nc - 1)
    c1 = np.polynomial.chebyshev.chebfit(t, np.sqrt(x)*scipy.special.k1e(x), nc - 1)
    c0 = np.ascontiguousarray(c0, dtype=np.float64)
    c1 = np.ascontiguousarray(c1, dtype=np.float64)
    lib().ipde_set_k0_table(hptr(c0), hptr(c1), nc)
    _k0_table_set.add(dev)


class StageTimer(object):
    """CUDA-event stage timer: mark(name) closes the stage `name`; results() -> {name: total ms}."""
    def __init__(self):
        self.events = []

    def mark(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream())
        self.events.append((name, e))

    def results(self):
        torch.cuda.synchronize()
        out = {}
        for (_, a), (name, b) in zip(self.events[:-1], self.events[1:]):
            if name != 'start':
                out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out


# ----------------------------------------------------------------------------------------
class FFT1D(object):
    def __init__(self, n):
        require_cuda()
        import ctypes
        self.n = n
        h = ctypes.c_void_p()
        lib().ipde_fft_plan_create(n, ctypes.byref(h))
        self.h = h

    def __del__(self):
        try:
            lib().ipde_fft_plan_destroy(self.h)
        except Exception:
            pass

    def c2c(self, x, sign=-1, scale=1.0, out=None):
        """x: complex128 (..., n) contiguous device tensor"""
        if out is None:
            out = torch.empty_like(x)
        batch = x.numel()//self.n
        lib().ipde_fft_c2c(self.h, ptr(x), ptr(out), batch, sign, scale, stream())
        return out

    def mul_fft(self, x, m, out=None):
        if out is None:
            out = torch.empty_like(x)
        lib().ipde_fft_mul_fft(self.h, ptr(x), ptr(out), ptr(m), x.numel()//self.n, stream())
        return out


class GridPlan(object):
    """Periodic (nx, ny) grid: 2-D real FFT pair, spectral multipliers, differences, interpolation."""
    def __init__(self, nx, ny, xh, yh):
        require_cuda()
        import ctypes
        self.nx, self.ny, self.xh, self.yh = nx, ny, xh, yh
        self.pitch = ny//2 + 1
        h = ctypes.c_void_p()
        lib().ipde_grid_plan_create(nx, ny, float(xh), float(yh), ctypes.byref(h))
        self.h = h
        n1, n2 = ctypes.c_int(), ctypes.c_int()
        lib().ipde_grid_plan_layout(h, ctypes.byref(n1), ctypes.byref(n2))
        self.n1, self.n2 = n1.value, n2.value

    def __del__(self):
        try:
            lib().ipde_grid_plan_destroy(self.h)
        except Exception:
            pass

    def row_of_freq(self):
        """row position of every frequency index (identity unless the column FFT is split)"""
        i = np.arange(self.nx)
        if self.n2 == 1:
            return i
        return self.n2*(i % self.n1) + i//self.n1

    def new_spec(self):
        return empty((self.nx, self.pitch), C128)

    def new_real(self):
        return empty((self.nx, self.ny), F64)

    def fft2(self, f, spec=None):
        if spec is None:
            spec = self.new_spec()
        lib().ipde_grid_fft2_r2c(self.h, ptr(f), ptr(spec), stream())
        return spec

    def ifft2(self, spec, work=None, out=None):
        if work is None:
            work = self.new_spec()
        if out is None:
            out = self.new_real()
        lib().ipde_grid_fft2_c2r(self.h, ptr(spec), ptr(work), ptr(out), stream())
        return out

    def spec_to_numpy_full(self, spec):
        """half spectrum (device, plan row order) -> full (nx, ny) complex numpy array in FFT order (tests)"""
        s = spec.cpu().numpy()[self.row_of_freq()]
        full = np.empty((self.nx, self.ny), dtype=complex)
        full[:, :self.pitch] = s
        j = np.arange(self.pitch, self.ny)
        i = (-np.arange(self.nx)) % self.nx
        full[:, self.pitch:] = np.conj(s[i][:, self.ny - j])
        return full


class Nufft2(object):
    """Type-2 NUFFT from a GridPlan half spectrum to fixed nonuniform points (angles in [0, 2 pi))."""
    def __init__(self, grid, x, y, eps=1e-14, sigma=2.0):
        import ctypes
        self.grid = grid
        self.w = es_kernel_width(eps)
        self.beta = 2.30*self.w
        self.fx = lib().ipde_fft_next_size(int(np.ceil(sigma*grid.nx)))
        self.fy = lib().ipde_fft_next_size(int(np.ceil(sigma*grid.ny)))
        px = np.ascontiguousarray(es_deconvolution(self.w, self.beta, grid.nx, self.fx))
        py = np.ascontiguousarray(es_deconvolution(self.w, self.beta, grid.ny, self.fy)[:grid.ny//2 + 1])
        h = ctypes.c_void_p()
        lib().ipde_nufft2_plan_create(grid.nx, grid.ny, float(grid.xh), float(grid.yh), self.fx, self.fy, self.w,
                                      self.beta, hptr(px), hptr(py), ctypes.byref(h))
        self.h = h
        self.x = to_dev(np.mod(x, 2*np.pi))
        self.y = to_dev(np.mod(y, 2*np.pi))
        self.npts = int(self.x.numel())

    def __del__(self):
        try:
            lib().ipde_nufft2_plan_destroy(self.h)
        except Exception:
            pass

    def __call__(self, spec, dx=0, dy=0, out=None):
        if out is None:
            out = empty((self.npts,))
        lib().ipde_nufft2_apply(self.h, ptr(spec), self.grid.n1, self.grid.n2, dx, dy, ptr(self.x), ptr(self.y),
                                self.npts, ptr(out), stream())
        return out


def _nufft2_apply3(self, spec):
    """(u, u_x, u_y) at the plan's points from one fine-grid transform"""
    v, dx, dy = empty((self.npts,)), empty((self.npts,)), empty((self.npts,))
    lib().ipde_nufft2_apply3(self.h, ptr(spec), self.grid.n1, self.grid.n2, ptr(self.x), ptr(self.y), self.npts,
                             ptr(v), ptr(dx), ptr(dy), stream())
    return v, dx, dy


Nufft2.apply3 = _nufft2_apply3


class RadialInterpolator(object):
    """(M, n) radial block -> values at points (theta, t): embedded_boundary.py:419-434 on the device."""
    def __init__(self, M, n, eps=1e-14, sigma=2.0):
        import ctypes
        self.M, self.n = M, n
        self.w = es_kernel_width(eps)
        self.beta = 2.30*self.w
        self.fn = lib().ipde_fft_next_size(int(np.ceil(sigma*n)))
        dec = np.ascontiguousarray(es_deconvolution(self.w, self.beta, n, self.fn))
        h = ctypes.c_void_p()
        lib().ipde_nufft1_plan_create(n, self.fn, self.w, self.beta, hptr(dec), ctypes.byref(h))
        self.h = h
        # DCT matrix: a_m = (2/M) sum_i f_i cos(m theta_i), a_0 halved; theta_i = (i + 1/2) pi / M
        th = (np.arange(M) + 0.5)*np.pi/M
        C = (2.0/M)*np.cos(np.outer(np.arange(M), th))
        C[0] *= 0.5
        self.cmat = to_dev(C)
        self.coef = empty((M, n))

    def __del__(self):
        try:
            lib().ipde_nufft1_plan_destroy(self.h)
        except Exception:
            pass

    def __call__(self, fr, theta, t, out=None, out_pos=None):
        npts = int(theta.numel())
        if out is None:
            out = empty((npts,))
        lib().ipde_radial_to_points(self.h, ptr(self.cmat), self.M, ptr(fr), ptr(self.coef), ptr(theta), ptr(t),
                                    npts, ptr(out), ptr(out_pos), stream())
        return out


def radial_to_points(fr, transf_r, t):
    """host convenience used by EmbeddedBoundary.interpolate_radial_to_points"""
    M, n = fr.shape
    R = RadialInterpolator(M, n)
    out = R(to_dev(fr), to_dev(transf_r), to_dev(np.mod(t, 2*np.pi)))
    return out.cpu().numpy()


class DenseLU(object):
    """LU factors (host scipy.linalg.lu_factor output) kept on the device for two-step solves."""
    def __init__(self, lu_piv):
        lu, piv = lu_piv
        M = lu.shape[0]
        perm = np.arange(M)
        for i, p in enumerate(piv):
            perm[i], perm[p] = perm[p], perm[i]
        self.M = M
        self.lu = to_dev(np.ascontiguousarray(lu))
        self.perm = to_dev(perm.astype(np.int32))

    def solve(self, b, out=None):
        if out is None:
            out = empty((self.M,))
        lib().ipde_lu_solve(ptr(self.lu), ptr(self.perm), ptr(b), ptr(out), self.M, stream())
        return out


def gemv(A, x, out=None):
    m, n = A.shape
    if out is None:
        out = empty((m,))
    lib().ipde_dense_gemv(ptr(A), ptr(x), ptr(out), m, n, stream())
    return out


def layer_apply_scalar(kind, k, sx, sy, q, tx, ty, out=None, accumulate=False):
    """kind: 0 Laplace, 1 modified Helmholtz; q already weighted"""
    if kind == 1:
        ensure_k0_table()
    nt = int(tx.numel())
    if out is None:
        out = empty((nt,))
    lib().ipde_layer_apply_scalar(kind, float(k), ptr(sx), ptr(sy), ptr(q), int(sx.numel()), ptr(tx), ptr(ty), nt,
                                  ptr(out), 1 if accumulate else 0, stream())
    return out


def layer_apply_scalar_dn(kind, k, sx, sy, q, tx, ty, tnx, tny, out=None, accumulate=False):
    """normal derivative at targets with unit normals (tnx, tny) of the sums of layer_apply_scalar"""
    if kind == 1:
        ensure_k0_table()
    nt = int(tx.numel())
    if out is None:
        out = empty((nt,))
    lib().ipde_layer_apply_scalar_dn(kind, float(k), ptr(sx), ptr(sy), ptr(q), int(sx.numel()), ptr(tx), ptr(ty),
                                     ptr(tnx), ptr(tny), nt, ptr(out), 1 if accumulate else 0, stream())
    return out


def layer_apply_stokes(sx, sy, fx, fy, tx, ty, out=None, accumulate=False):
    nt = int(tx.numel())
    if out is None:
        out = empty((3, nt))
    lib().ipde_layer_apply_stokes(ptr(sx), ptr(sy), ptr(fx), ptr(fy), int(sx.numel()), ptr(tx), ptr(ty), nt,
                                  ptr(out), 1 if accumulate else 0, stream())
    return out


# ----------------------------------------------------------------------------------------
class DeviceCollection(object):
    """
    Device-resident copy of what the per-solve path reads from a registered
    EmbeddedBoundaryCollection (ipde/ebdy_collection.py:351-526): the compact<->grid index
    lists, the cutoff, interface and radial point sets, per-boundary annulus coordinates.
    """
    def __init__(self, ebdyc):
        require_cuda()
        g = ebdyc.grid
        self.nx, self.ny = g.Nx, g.Ny
        self.grid = GridPlan(g.Nx, g.Ny, g.xh, g.yh)
        phys = ebdyc.phys
        lin = np.flatnonzero(phys.ravel())
        self.n_phys = lin.size
        self.phys_lin = to_dev(lin.astype(np.int64))
        self.step_compact = to_dev(ebdyc.grid_step[phys])
        # positions (in the compact physical vector) of the phys-not-in-annulus points
        self.pna_pos = to_dev(ebdyc.phys_inds[ebdyc.phys_not_in_annulus].astype(np.int64))
        self.n_pna = int(self.pna_pos.numel())
        self._ebdyc = ebdyc
        self._pnai = None                 # (x, y) of grid_pnai on the device: only the all-pairs back-end reads them
        self.n_pnai = ebdyc.grid_pnai.N
        self.iv_x = to_dev(ebdyc.all_ivx)
        self.iv_y = to_dev(ebdyc.all_ivy)
        self.n_iv = ebdyc.all_ivx.size
        self.bumpy = to_dev(ebdyc.bumpy) if getattr(ebdyc, 'bumpy_readied', False) else None
        self.x0, self.y0 = float(g.xv[0]), float(g.yv[0])
        self._nufft = None
        self._ebdyc_transf = (ebdyc.interfaces_x_transf, ebdyc.interfaces_y_transf)
        self.radial_offsets = np.concatenate([[0], np.cumsum(ebdyc.radial_dof_list)]) + self.n_phys
        self.bdy_offsets = np.concatenate([[0], np.cumsum(ebdyc.bdy_Ns)])
        self.ebdys = [DeviceBoundary(e) for e in ebdyc]
        # work buffers for the grid solve
        self.fc = self.grid.new_real()
        self.uc = self.grid.new_real()
        self.uch = self.grid.new_spec()
        self.spec_work = self.grid.new_spec()

    @property
    def pnai_x(self):
        if self._pnai is None:
            self._pnai = (to_dev(self._ebdyc.grid_pnai.x), to_dev(self._ebdyc.grid_pnai.y))
        return self._pnai[0]

    @property
    def pnai_y(self):
        self.pnai_x
        return self._pnai[1]

    @property
    def nufft(self):
        if self._nufft is None:
            self._nufft = Nufft2(self.grid, self._ebdyc_transf[0], self._ebdyc_transf[1])
        return self._nufft

    def interpolate_grid_to_interface(self, f_host, order):
        """host-array convenience behind ebdyc.interpolate_grid_to_interface (:578-614)"""
        f = to_dev(np.asarray(f_host, dtype=float))
        if order == np.inf:
            spec = self.grid.fft2(f)
            return self.nufft(spec).cpu().numpy()
        if order != 3:
            raise IpdeError('device polynomial interpolation implements order 3 (solver_type="fourth") only')
        out = empty((self.n_iv,))
        lib().ipde_grid_interp_cubic(self.grid.h, ptr(f), self.x0, self.y0, ptr(self.iv_x), ptr(self.iv_y),
                                     self.n_iv, ptr(out), stream())
        return out.cpu().numpy()


class DeviceBoundary(object):
    def __init__(self, ebdy):
        self.M, self.n = ebdy.M, ebdy.bdy.N
        self.interior = ebdy.interior
        # in-annulus grid points that the inside test calls exterior (within round-off of the boundary) have no slot
        # in the compact vector: drop them (the reference writes them into the full grid and masks them afterwards,
        # multi_boundary/scalar.py:110-114)
        ok = np.asarray(getattr(ebdy, 'ia_is_phys', np.ones(ebdy.ia_inds.shape, dtype=bool)), dtype=bool)
        self.ia_theta = to_dev(np.ascontiguousarray(ebdy.grid_ia_r_transf[ok]))
        self.ia_t = to_dev(np.ascontiguousarray(np.mod(ebdy.grid_ia_t, 2*np.pi)[ok]))
        self.ia_pos = to_dev(np.ascontiguousarray(np.asarray(ebdy.ia_inds, dtype=np.int64)[ok]))
        self.n_ia = int(self.ia_pos.numel())
        self.radial_x = to_dev(ebdy.radial_x.ravel())
        self.radial_y = to_dev(ebdy.radial_y.ravel())
        self.interface_x = to_dev(ebdy.interface.x)
        self.interface_y = to_dev(ebdy.interface.y)
        self.interface_nx = to_dev(ebdy.interface.normal_x)
        self.interface_ny = to_dev(ebdy.interface.normal_y)
        self.radial_interp = RadialInterpolator(self.M, self.n)


# ----------------------------------------------------------------------------------------
class FreeSpaceEvaluator(object):
    """
    Spectral-Ewald evaluation of  u(x) = sum_s G(|x - y_s|) q_s  (csrc/ewald.cu, ipde_b200/ewald.py) from fixed
    sources to fixed targets: nodes of the background grid (`grid_nodes`: (ix, iy) index arrays, results are
    *added* to out[grid_out_pos]) and/or off-grid points (`points`: x, y arrays).
      kind 0: G = -log r / (2 pi);  kind 1: G = K0(kappa r) / (2 pi)
      grid: object with xv[0], yv[0], xh (== yh) -- FFT-box nodes are aligned with the background grid
    """
    def __init__(self, kind, kappa, grid_x0, grid_y0, h, sx, sy, grid_nodes=None, grid_out_pos=None, points=None):
        import ctypes
        from . import ewald as E
        require_cuda()
        L_ = lib()
        self.kind, self.kappa, self.h = int(kind), float(kappa), float(h)
        P = E.EwaldParameters(h)
        self.params = P
        sx = np.ascontiguousarray(sx, dtype=float); sy = np.ascontiguousarray(sy, dtype=float)
        self.nsrc = sx.size
        # extents and the largest distance from the centre over sources and targets; the grid targets (1e8 of them at
        # 16384^2) are handled as index tensors on the device
        xs, ys = [sx], [sy]
        if points is not None:
            ptx = np.ascontiguousarray(points[0], dtype=float); pty = np.ascontiguousarray(points[1], dtype=float)
            xs.append(ptx); ys.append(pty)
        lox, hix = min(a.min() for a in xs), max(a.max() for a in xs)
        loy, hiy = min(a.min() for a in ys), max(a.max() for a in ys)
        if grid_nodes is not None:
            gix = to_dev(np.asarray(grid_nodes[0]), torch.int64)
            giy = to_dev(np.asarray(grid_nodes[1]), torch.int64)
            lox, hix = min(lox, grid_x0 + h*int(gix.min())), max(hix, grid_x0 + h*int(gix.max()))
            loy, hiy = min(loy, grid_y0 + h*int(giy.min())), max(hiy, grid_y0 + h*int(giy.max()))
        cx, cy = 0.5*(lox + hix), 0.5*(loy + hiy)
        r2max = max(((a - cx)**2 + (b - cy)**2).max() for a, b in zip(xs, ys))
        if grid_nodes is not None:
            r2max = max(r2max, float(((grid_x0 + h*gix - cx)**2 + (grid_y0 + h*giy - cy)**2).max()))
        D = min(np.hypot(hix - lox, hiy - loy), 2.0*np.sqrt(r2max))
        ax, ay = np.array([lox, hix]), np.array([loy, hiy])
        self.L = D + 2.0*np.sqrt(2.0)*P.m*h + 2.0*h
        nxt = L_.ipde_fft_next_size
        self.x0, self.px, self.ilo_x = E.choose_box(ax.min(), ax.max(), h, P.m, grid_x0, self.L, nxt)
        self.y0, self.py, self.ilo_y = E.choose_box(ay.min(), ay.max(), h, P.m, grid_y0, self.L, nxt)
        tab = E.local_tables(self.kind, self.kappa, P.xi, P.rc)
        tix, tiy, tptr, tsrc, tmap = E.build_tiles(sx, sy, self.x0, self.y0, h, self.px, self.py, P.rc)
        c0, c1, _ = E.trunc_kernel_constants(self.kind, self.kappa, self.L)
        self.ntiles = int(tix.size)
        d_sx, d_sy = to_dev(sx), to_dev(sy)
        hnd = ctypes.c_void_p()
        tS, tA = np.ascontiguousarray(tab['S']), np.ascontiguousarray(tab['A'])
        tmap_c = np.ascontiguousarray(tmap, dtype=np.int32)
        L_.ipde_ewald_plan_create(self.kind, self.kappa, float(P.xi), self.h, float(P.rc), int(P.m), float(self.x0),
                                  float(self.y0), int(self.px), int(self.py), float(self.L), float(c0), float(c1),
                                  self.nsrc, ptr(d_sx), ptr(d_sy), tab['nint'], tab['ndeg'], tab['nlog'], hptr(tS),
                                  hptr(tA), self.ntiles, hptr(tix), hptr(tiy), hptr(tptr), hptr(tsrc), hptr(tmap_c),
                                  ctypes.byref(hnd))
        self.h_plan = hnd
        self._tab = tab
        self._grad_ready = False
        self.xc, self.yc = float(cx), float(cy)         # origin of the affine weights of apply_fields
        self.d_sx, self.d_sy = d_sx, d_sy
        self.n_grid = 0
        self.n_points = 0
        if grid_nodes is not None:
            bix, biy = gix - self.ilo_x, giy - self.ilo_y          # node indices inside the FFT box
            d_lin = bix*self.py + biy
            t = to_dev(tmap_c, torch.int64)[bix//E.TILE, biy//E.TILE]
            d_loc = torch.where(t >= 0, t*(E.TILE*E.TILE) + (bix % E.TILE)*E.TILE + (biy % E.TILE),
                                torch.full_like(t, -1))
            d_pos = to_dev(np.asarray(grid_out_pos), torch.int64)
            self.n_grid = int(d_lin.numel())
            del gix, giy, bix, biy, t
            L_.ipde_ewald_set_grid_targets(hnd, self.n_grid, ptr(d_lin), ptr(d_loc), ptr(d_pos))
        if points is not None:
            self.n_points = int(ptx.size)
            d_ptx, d_pty = to_dev(ptx), to_dev(pty)
            L_.ipde_ewald_set_point_targets(hnd, self.n_points, ptr(d_ptx), ptr(d_pty))
        torch.cuda.synchronize()

    def __del__(self):
        try:
            lib().ipde_ewald_plan_destroy(self.h_plan)
        except Exception:
            pass

    def apply_fields(self, q, grid_terms=(), point_terms=()):
        """
        Value and gradient with affine weights: every term is (field, dst, (w0, wx, wy)) with field 0 value,
        1 d/dx, 2 d/dy; dst[.] += (w0 + wx (x - xc) + wy (y - yc)) field  at the grid targets (dst indexed by
        grid_out_pos) or at the point targets (dst indexed by point).  At most two terms per field and target set.
        """
        L_ = lib()
        if not self._grad_ready:
            dS, dA = np.ascontiguousarray(self._tab['dS']), np.ascontiguousarray(self._tab['dA'])
            L_.ipde_ewald_set_gradient_tables(self.h_plan, hptr(dS), hptr(dA))
            self._grad_ready = True
        def pack(terms):
            w = np.zeros((3, 2, 3))
            d = np.zeros((3, 2), dtype=np.int64)
            used = [0, 0, 0]
            for field, dst, wt in terms:
                k = used[field]
                if k >= 2:
                    raise IpdeError('at most two destinations per field')
                w[field, k] = wt
                d[field, k] = dst.data_ptr()
                used[field] += 1
            return np.ascontiguousarray(w), np.ascontiguousarray(d)
        gw, gd = pack(grid_terms)
        pw, pd = pack(point_terms)
        L_.ipde_ewald_apply_fields(self.h_plan, ptr(q), self.xc, self.yc, hptr(gw), hptr(gd), hptr(pw), hptr(pd), stream())

    def __call__(self, q, grid_out=None, point_out=None, accumulate_points=False):
        """q: weighted charges (device).  grid_out[grid_out_pos] += u; returns point values (device) or None."""
        if self.n_points > 0 and point_out is None:
            point_out = empty((self.n_points,))
        lib().ipde_ewald_apply(self.h_plan, ptr(q), ptr(grid_out), ptr(point_out), 1 if accumulate_points else 0,
                               stream())
        return point_out


class StokesFreeSpaceEvaluator(object):
    """
    Stokeslet velocity and pressure sums (the pyfmmlib2d.SFMM call of ipde/solvers/internals/stokes.py:26-35) through
    the Laplace spectral-Ewald evaluator:  with phi_j = sum G_L f_j, chi = sum G_L (y_s . f_s), G_L = -log r / 2 pi,
        u_i = (phi_i - x_j d_i phi_j + d_i chi) / 2 ,      p = -d_j phi_j
    (x, y_s measured from the centre of the evaluator box).  Three charge sets, each one spread + forward transform;
    8 inverse field passes for grid targets, 3 for point targets.
    """
    def __init__(self, grid_x0, grid_y0, h, sx, sy, grid_nodes=None, grid_out_pos=None, points=None):
        self.ev = FreeSpaceEvaluator(0, 0.0, grid_x0, grid_y0, h, sx, sy, grid_nodes=grid_nodes,
                                     grid_out_pos=grid_out_pos, points=points)
        e = self.ev
        self.sxc = to_dev(np.ascontiguousarray(sx, dtype=float) - e.xc)
        self.syc = to_dev(np.ascontiguousarray(sy, dtype=float) - e.yc)
        self.n_points = e.n_points
        self._chi = empty((e.nsrc,))
        self._tmp = empty((e.nsrc,))

    def __call__(self, fx, fy, grid_out=None, point_out=None):
        """fx, fy: weighted force components (device, nsrc).  grid_out: (u, v, p) device vectors indexed by
        grid_out_pos (accumulated into); point_out: (3, npts) device array (accumulated into; allocated zeroed if
        None and the evaluator has point targets).  Returns point_out."""
        e = self.ev
        L_ = lib()
        n = e.nsrc
        L_.ipde_vec_mul(ptr(self.sxc), ptr(fx), ptr(self._chi), n, stream())
        L_.ipde_vec_mul(ptr(self.syc), ptr(fy), ptr(self._tmp), n, stream())
        L_.ipde_vec_axpby(1.0, ptr(self._chi), 1.0, ptr(self._tmp), ptr(self._chi), n, stream())
        if e.n_points > 0 and point_out is None:
            point_out = zeros((3, e.n_points))          # a tuple of three device vectors is accepted as well
        def terms(dst):
            if dst is None:
                return (), (), ()
            u, v, p = dst
            t1 = [(0, u, (0.5, 0, 0)), (1, u, (0, -0.5, 0)), (1, p, (-1.0, 0, 0)), (2, v, (0, -0.5, 0))]
            t2 = [(0, v, (0.5, 0, 0)), (1, u, (0, 0, -0.5)), (2, v, (0, 0, -0.5)), (2, p, (-1.0, 0, 0))]
            t3 = [(1, u, (0.5, 0, 0)), (2, v, (0.5, 0, 0))]
            return t1, t2, t3
        g1, g2, g3 = terms(grid_out if e.n_grid > 0 else None)
        p1, p2, p3 = terms((point_out[0], point_out[1], point_out[2]) if point_out is not None else None)
        e.apply_fields(fx, g1, p1)
        e.apply_fields(fy, g2, p2)
        e.apply_fields(self._chi, g3, p3)
        return point_out
