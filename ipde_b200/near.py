"""
Setup-time search: which grid points lie within a distance d of a closed curve,
what are their normal coordinates (r, t) with x = X(t) + r n(t), and which grid
points are inside the curve.

The reference takes this from the un-vendored `near_finder` package
(`gridpoints_near_curve_update`, `points_inside_curve_update`,
`compute_local_coordinates`; call sites ipde/embedded_boundary.py:11-12,202-206 and
ipde/ebdy_collection.py:11-13,113-122,380).  Its results are pinned only by the identity
x = X(t) + r n(t) (embedded_boundary.py:299-300), which tests/test_geometry.py checks
to 1e-13.  This restatement marks candidate points by scanline windows around the
boundary nodes, gets an initial guess from a k-d tree and runs Newton on
d/dt |X(t) - x|^2 = 0, evaluating the trigonometric interpolant X(t) and its
derivatives through local Taylor expansions about a 2x-upsampled set of nodes.

Host NumPy only; nothing here is on the per-solve path.
"""
import numpy as np
from scipy.spatial import cKDTree
from .geometry import fourier_resample

TAYLOR_ORDER = 16


class CurveTaylor(object):
    """Taylor tables of the trigonometric interpolant of a closed curve."""
    def __init__(self, c, upsample=2, order=TAYLOR_ORDER):
        c = np.asarray(c, dtype=complex)
        N = c.size
        self.N = N
        self.nc = upsample*N
        self.dtc = 2.0*np.pi/self.nc
        # upsample first (the coarse Nyquist mode becomes an interior mode pair),
        # then differentiate spectrally on the fine grid
        ch = np.fft.fft(fourier_resample(c, self.nc))
        k = np.fft.fftfreq(self.nc, 1.0/self.nc)
        self.order = order
        tab = np.empty((order + 1, self.nc), dtype=complex)
        fact = 1.0
        for m in range(order + 1):
            if m > 0:
                fact *= m
            tab[m] = np.fft.ifft(ch*(1j*k)**m)/fact
        self.tab = tab
        self.tc = np.arange(self.nc)*self.dtc

    def eval(self, ci, s):
        """X, X', X'' at t = tc[ci] + s  (ci int array, s float array)"""
        tab = self.tab
        P = self.order
        X = tab[P][ci].copy()
        Xp = np.zeros_like(X)
        Xpp = np.zeros_like(X)
        for m in range(P - 1, -1, -1):
            Xpp = Xpp*s + 2.0*Xp
            Xp = Xp*s + X
            X = X*s + tab[m][ci]
        return X, Xp, Xpp


def compute_local_coordinates(ct, z, guess_ci, tol=1e-14, maxiter=50):
    """
    Newton solve for the closest-point parameter.
    ct: CurveTaylor;  z: complex targets;  guess_ci: index of a nearby expansion centre.
    Returns t in [0, 2pi), r (signed distance along the outward normal of a CCW curve).
    """
    ci = np.asarray(guess_ci, dtype=np.int64) % ct.nc
    s = np.zeros(z.shape, dtype=float)
    active = np.ones(z.shape, dtype=bool)
    for it in range(maxiter):
        idx = np.nonzero(active)[0]
        if idx.size == 0:
            break
        X, Xp, Xpp = ct.eval(ci[idx], s[idx])
        d = X - z[idx]
        f = (d*np.conj(Xp)).real
        fp = (Xp*np.conj(Xp)).real + (d*np.conj(Xpp)).real
        # guard against non-convex steps far from the curve
        fp = np.where(fp <= 0.0, (Xp*np.conj(Xp)).real, fp)
        ds = -f/fp
        ds = np.clip(ds, -4*ct.dtc, 4*ct.dtc)
        snew = s[idx] + ds
        # re-centre so |s| <= dtc/2
        shift = np.rint(snew/ct.dtc).astype(np.int64)
        ci[idx] = (ci[idx] + shift) % ct.nc
        s[idx] = snew - shift*ct.dtc
        active[idx] = np.abs(ds) > tol
    X, Xp, _ = ct.eval(ci, s)
    sp = np.abs(Xp)
    n = -1j*Xp/sp
    r = ((z - X)*np.conj(n)).real
    t = (ct.tc[ci] + s) % (2.0*np.pi)
    return t, r


class NearCurveResult(object):
    """Grid points within distance d of the curve with their (r, t) coordinates."""
    def __init__(self, ix, iy, r, t):
        self.ix, self.iy, self.r, self.t = ix, iy, r, t
        self.n = ix.size

    # tuple-like access in the order the reference unpacks it
    # (embedded_boundary.py:211-215: nclose, iax, iay, r, t)
    def __getitem__(self, i):
        return (self.n, self.ix, self.iy, self.r, self.t)[i]


# grids with at least this many points are registered on the GPU when there is one (torch tensor ops: run-length
# marking + prefix sums + compaction for the near-curve search and the inside test, Taylor-table Newton for the
# coordinates); IPDE_REGISTER=host forces the NumPy path
DEVICE_MIN_POINTS = 1 << 22


def _torch_device(npoints):
    import os
    if os.environ.get('IPDE_REGISTER', '') == 'host' or npoints < DEVICE_MIN_POINTS:
        return None
    try:
        import torch
        return torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else None
    except Exception:
        return None


def _mark_runs_device(dev, rows, lo, hi, ok, Nx, Ny):
    """(ix, iy) of all grid points covered by at least one run [lo, hi) of row `rows` -- on the device, restricted to
    the bounding box of the runs; same row-major order as np.nonzero"""
    import torch
    r = rows[ok]; a = lo[ok]; b = hi[ok]
    if r.size == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z
    r0, r1 = int(r.min()), int(r.max()) + 1
    c0, c1 = int(a.min()), int(b.max())
    diff = torch.zeros((r1 - r0, c1 - c0 + 1), dtype=torch.int32, device=dev)
    tr = torch.from_numpy(r - r0).to(dev)
    one = torch.ones(tr.numel(), dtype=torch.int32, device=dev)
    diff.index_put_((tr, torch.from_numpy(a - c0).to(dev)), one, accumulate=True)
    diff.index_put_((tr, torch.from_numpy(b - c0).to(dev)), -one, accumulate=True)
    close = torch.cumsum(diff, dim=1, dtype=torch.int32)[:, :c1 - c0] > 0
    idx = torch.nonzero(close)
    return (idx[:, 0] + r0).cpu().numpy(), (idx[:, 1] + c0).cpu().numpy()


def _local_coordinates_device(dev, ct, z, guess_ci, tol=1e-14, maxiter=50):
    """compute_local_coordinates with torch tensors on `dev` (same iteration, all points advanced together)"""
    import torch
    tab = torch.from_numpy(ct.tab).to(dev)                     # (order+1, nc) complex
    P, nc, dtc = ct.order, ct.nc, ct.dtc
    zt = torch.from_numpy(z).to(dev)
    ci = torch.from_numpy(np.asarray(guess_ci, dtype=np.int64) % nc).to(dev)
    s = torch.zeros(zt.shape, dtype=torch.float64, device=dev)

    def ev(ci_, s_):
        X = tab[P][ci_].clone()
        Xp = torch.zeros_like(X)
        Xpp = torch.zeros_like(X)
        for m in range(P - 1, -1, -1):
            Xpp = Xpp*s_ + 2.0*Xp
            Xp = Xp*s_ + X
            X = X*s_ + tab[m][ci_]
        return X, Xp, Xpp
    active = torch.ones(zt.shape, dtype=torch.bool, device=dev)
    for it in range(maxiter):
        idx = torch.nonzero(active)[:, 0]
        if idx.numel() == 0:
            break
        X, Xp, Xpp = ev(ci[idx], s[idx])
        dd = X - zt[idx]
        f = (dd*Xp.conj()).real
        sp2 = (Xp*Xp.conj()).real
        fp = sp2 + (dd*Xpp.conj()).real
        fp = torch.where(fp <= 0.0, sp2, fp)
        ds = torch.clamp(-f/fp, -4*dtc, 4*dtc)
        snew = s[idx] + ds
        shift = torch.round(snew/dtc).to(torch.int64)
        ci[idx] = (ci[idx] + shift) % nc
        s[idx] = snew - shift*dtc
        active[idx] = ds.abs() > tol
    X, Xp, _ = ev(ci, s)
    n = -1j*Xp/Xp.abs()
    r = ((zt - X)*n.conj()).real
    t = (torch.from_numpy(ct.tc).to(dev)[ci] + s) % (2.0*np.pi)
    return t.cpu().numpy(), r.cpu().numpy()


def gridpoints_near_curve(bdy, xv, yv, d, tol=1e-14):
    """
    All grid points (xv[i], yv[j]) whose distance to the curve is <= d.
    """
    xh = xv[1] - xv[0]
    yh = yv[1] - yv[0]
    Nx, Ny = xv.size, yv.size
    # polygon fine enough that every point within d of the curve is within
    # sqrt(d^2 + (h/2)^2) of a node
    up = max(1, int(np.ceil(bdy.max_h/min(xh, yh))))
    cf = fourier_resample(bdy.c, up*bdy.N) if up > 1 else bdy.c
    R = d + 0.75*max(xh, yh)
    # scanline marking: for each node and each row offset, a run of columns
    wx = int(np.ceil(R/xh)) + 1
    cx = np.rint((cf.real - xv[0])/xh).astype(np.int64)
    offs = np.arange(-wx, wx + 1)
    rows = cx[:, None] + offs[None, :]
    dxr = (xv[0] + rows*xh) - cf.real[:, None]
    hw2 = R*R - dxr*dxr
    ok = np.logical_and(hw2 >= 0.0, np.logical_and(rows >= 0, rows < Nx))
    hw = np.sqrt(np.where(hw2 > 0, hw2, 0.0))
    lo = np.ceil((cf.imag[:, None] - hw - yv[0])/yh).astype(np.int64)
    hi = np.floor((cf.imag[:, None] + hw - yv[0])/yh).astype(np.int64)
    lo = np.clip(lo, 0, Ny)
    hi = np.clip(hi + 1, 0, Ny)
    ok = np.logical_and(ok, hi > lo)
    dev = _torch_device(Nx*Ny)
    if dev is not None:
        ix, iy = _mark_runs_device(dev, rows, lo, hi, ok, Nx, Ny)
    else:
        diff = np.zeros((Nx, Ny + 1), dtype=np.int32)
        np.add.at(diff, (rows[ok], lo[ok]), 1)
        np.add.at(diff, (rows[ok], hi[ok]), -1)
        close = np.cumsum(diff, axis=1)[:, :Ny] > 0
        ix, iy = np.nonzero(close)
    if ix.size == 0:
        e = np.zeros(0)
        return NearCurveResult(ix, iy, e, e)
    z = xv[ix] + 1j*yv[iy]
    ct = CurveTaylor(bdy.c)
    tree = cKDTree(np.column_stack([ct.tab[0].real, ct.tab[0].imag]))
    _, gi = tree.query(np.column_stack([z.real, z.imag]), workers=-1)
    if dev is not None:
        t, r = _local_coordinates_device(dev, ct, z, gi, tol=tol)
    else:
        t, r = compute_local_coordinates(ct, z, gi, tol=tol)
    keep = np.abs(r) <= d
    return NearCurveResult(ix[keep], iy[keep], r[keep], t[keep])


def points_inside_curve(bdy, xv, yv, near=None):
    """
    Boolean (Nx, Ny): grid point strictly inside the closed curve.  Parity fill
    along y for every grid line x = xv[i], using a fine polygon; points in
    `near` (a NearCurveResult) are then classified exactly by the sign of r.
    """
    xh = xv[1] - xv[0]
    yh = yv[1] - yv[0]
    Nx, Ny = xv.size, yv.size
    up = max(4, int(np.ceil(4.0*bdy.max_h/min(xh, yh))))
    cf = fourier_resample(bdy.c, up*bdy.N)
    x0, y0 = cf.real, cf.imag
    x1, y1 = np.roll(x0, -1), np.roll(y0, -1)
    # grid lines crossed by each edge (half-open rule on x)
    a = (np.minimum(x0, x1) - xv[0])/xh
    b = (np.maximum(x0, x1) - xv[0])/xh
    ilo = np.ceil(a).astype(np.int64)
    ihi = np.ceil(b).astype(np.int64) - 1
    cnt = np.maximum(ihi - ilo + 1, 0)
    cross_i, cross_j = [], []
    for m in range(int(cnt.max()) if cnt.size else 0):
        sel = cnt > m
        i = ilo[sel] + m
        good = np.logical_and(i >= 0, i < Nx)
        xs = xv[0] + i*xh
        yc = y0[sel] + (xs - x0[sel])/(x1[sel] - x0[sel])*(y1[sel] - y0[sel])
        j = np.floor((yc - yv[0])/yh).astype(np.int64) + 1
        j = np.clip(j, 0, Ny)
        cross_i.append(i[good]); cross_j.append(j[good])
    ci_ = np.concatenate(cross_i) if cross_i else np.zeros(0, dtype=np.int64)
    cj_ = np.concatenate(cross_j) if cross_j else np.zeros(0, dtype=np.int64)
    dev = _torch_device(Nx*Ny)
    if dev is not None and ci_.size:
        # parity fill on the device, restricted to the rows / columns the curve spans (outside: not inside)
        import torch
        r0, r1 = int(ci_.min()), int(ci_.max()) + 1
        c0 = int(cj_.min())
        diff = torch.zeros((r1 - r0, Ny + 1 - c0), dtype=torch.int32, device=dev)
        diff.index_put_((torch.from_numpy(ci_ - r0).to(dev), torch.from_numpy(cj_ - c0).to(dev)),
                        torch.ones(ci_.size, dtype=torch.int32, device=dev), accumulate=True)
        sub = (torch.cumsum(diff, dim=1, dtype=torch.int32)[:, :Ny - c0] % 2) == 1
        inside = np.zeros((Nx, Ny), dtype=bool)
        inside[r0:r1, c0:] = sub.cpu().numpy()
    else:
        diff = np.zeros((Nx, Ny + 1), dtype=np.int32)
        np.add.at(diff, (ci_, cj_), 1)
        inside = (np.cumsum(diff, axis=1)[:, :Ny] % 2) == 1
    if near is not None and near.n > 0:
        inside[near.ix, near.iy] = near.r < 0.0
    return inside
