"""
Spectral-Ewald free-space evaluator: host-side parameters, tables and geometry preprocessing.

What it replaces.  The reference evaluates  u(x) = sum_s G(|x - y_s|) q_s  from all effective (QFS) sources
to (a) every physical grid point outside the annuli and every interface point
(`ScalarSolver.evaluate_to_grid_pnai`, ipde/solvers/multi_boundary/scalar.py:63-71) and (b) the radial points
of each annulus (`ScalarHelper.correct`, ipde/solvers/internals/scalar.py:95-116), either by a direct
O(S T) sum / FMM (`Layer_Apply`) or with its "Ewald-like" grid evaluator
(ipde/grid_evaluators/scalar_grid_evaluator.py:188-231,266-307: local screened kernel spread around each
source + one FFT solve with the truncated spectral Green's function on a zero-padded grid).  The reference's
evaluator is limited to ~10 digits by a fitted residual operator (scalar_grid_evaluator.py:60-63); here the
split is the classical Gaussian (Ewald) one, which is exact to the chosen tolerance:

    G = G_loc + G_far,   G_far = G * gamma_xi,   gamma_xi = (xi^2/pi) exp(-xi^2 r^2)
    G_loc(r) = (1/4pi) int_1^inf exp(-xi^2 r^2 u - kappa^2/(4 xi^2 u)) du/u     (kappa = 0: E1(xi^2 r^2)/4pi)

G_loc is summed directly over the sources within rc = sw*h of a target.  The far part is computed on the
uniform grid with the spectral-Ewald factorisation gamma_xi = w * w, w = (2 xi^2/pi) exp(-2 xi^2 r^2)
(Lindbo & Tornberg 2011): spread the charges with w, FFT, multiply by the truncated kernel
(Vico, Greengard & Ferrando 2016; same `trunc_sgf` formulas as laplace_grid_evaluator.py:21-33 and
modified_helmholtz_grid_evaluator.py:14-17), inverse FFT:

    H = ifft2( G_T^(k) fft2(S_w) )                          gathered with w at off-grid targets
    U = ifft2( G_T^(k) exp(-k^2/8 xi^2) fft2(S_w) )         the far field itself on the grid nodes

The FFT box has period >= (extent of sources and targets) + L with L >= the largest source-target
distance, so the radially truncated kernel can be sampled directly on the FFT grid (no 4x precomputation
grid as in scalar_grid_evaluator.py:277-293).

This module holds the NumPy side: parameters, the truncated kernel, the piecewise-polynomial tables of
G_loc uploaded to the device, tile/source lists, and `eval_local` (NumPy evaluation of the same tables, used
by the tests).  The CUDA side is ipde_b200/csrc/ewald.cu.
"""
import numpy as np
import scipy.special as ss

TILE = 16
EULER = 0.57721566490153286061
NINT = 32          # intervals in z = r^2 on [0, rc^2]
NDEG = 14          # coefficients per interval (degree 13)


class EwaldParameters(object):
    """xi, local cutoff rc, window cutoff rw (all tied to the grid spacing h) for a tolerance ~1e-14."""
    def __init__(self, h, sw=21):
        self.h = float(h)
        self.sw = int(sw)
        self.rc = self.sw*self.h
        self.xi = np.sqrt(np.log(1.0e14))/self.rc          # exp(-xi^2 rc^2) = 1e-14 ; xi*h = 0.27
        self.m = int(np.ceil(self.sw/np.sqrt(2.0)))        # window half width in cells (15)
        self.rw = self.m*self.h


def trunc_kernel_hat(kind, kappa, L, k):
    """Fourier transform of G(r) 1_{r<L} at |k| (G = -log r/2pi for kind 0, K0(kappa r)/2pi for kind 1)."""
    k = np.asarray(k, dtype=float)
    if kind == 0:
        z = k == 0
        kn = np.where(z, 1.0, k)
        out = (1.0 - ss.j0(L*kn))/kn**2 - L*np.log(L)*ss.j1(L*kn)/kn
        return np.where(z, L*L/4 - L*L*np.log(L)/2, out)
    return (1.0 + L*k*ss.j1(L*k)*ss.k0(L*kappa) - L*kappa*ss.j0(L*k)*ss.k1(L*kappa))/(k**2 + kappa**2)


def trunc_kernel_constants(kind, kappa, L):
    """(c0, c1, c2) such that the device evaluates the formula above as
       kind 0: (1 - J0(Lk))/k^2 - c0 J1(Lk)/k, value c1 at k = 0
       kind 1: (1 + c0 k J1(Lk) - c1 J0(Lk))/(k^2 + kappa^2)"""
    if kind == 0:
        return L*np.log(L), L*L/4 - L*L*np.log(L)/2, 0.0
    return L*ss.k0(L*kappa), L*kappa*ss.k1(L*kappa), 0.0


def _gloc(kind, kappa, xi, r):
    """G_loc(r) by quadrature of (1/4pi) int_1^inf exp(-a u - b/u) du/u, a = xi^2 r^2, b = kappa^2/4xi^2"""
    r = np.atleast_1d(np.asarray(r, dtype=float))
    a = (xi*r)**2
    if kind == 0:
        return ss.exp1(a)/(4*np.pi)
    b = kappa*kappa/(4*xi*xi)
    x, w = np.polynomial.legendre.leggauss(300)
    out = np.empty_like(r)
    for i, ai in enumerate(a):
        umax = 1.0 + 45.0/ai
        # u = 1 + (umax - 1) t^2 concentrates nodes near u = 1 where the integrand is largest
        t = 0.5*(x + 1.0)
        u = 1.0 + (umax - 1.0)*t*t
        du = (umax - 1.0)*2.0*t*0.5*w
        out[i] = np.sum(np.exp(-ai*u - b/u)/u*du)/(4*np.pi)
    return out


def _gfar(kappa, xi, r):
    """modified Helmholtz G_far(r) = (1/4pi) int_0^1 exp(-xi^2 r^2 u - kappa^2/(4 xi^2 u)) du/u (smooth in r^2)"""
    r = np.atleast_1d(np.asarray(r, dtype=float))
    b = kappa*kappa/(4*xi*xi)
    # s = -log u in [0, smax]: integrand exp(-a e^-s - b e^s) ds
    smax = np.log(45.0/b) if b < 45.0 else 0.0
    if smax <= 0.0:
        return np.zeros_like(r)
    x, w = np.polynomial.legendre.leggauss(400)
    s = 0.5*smax*(x + 1.0)
    ws = 0.5*smax*w
    a = (xi*r)**2
    return np.array([np.sum(np.exp(-ai*np.exp(-s) - b*np.exp(s))*ws) for ai in a])/(4*np.pi)


NLOG = 4           # intervals next to r = 0 on which the log singularity is split off


def local_tables(kind, kappa, xi, rc, nint=NINT, ndeg=NDEG, nlog=NLOG, check=True):
    """
    Piecewise polynomials in z = r^2 on [0, rc^2]:  G_loc(r) = S_i(t) - A_i(t) log(z),  t in [-1, 1] the local
    variable of interval i = floor(nint z / rc^2).  A_i is nonzero only on the first `nlog` intervals (log singularity
    at r = 0, which also limits the convergence of a direct fit on the intervals next to it);
    elsewhere G_loc is tabulated directly.  Returns dict with 'S', 'A' (nint, ndeg) Chebyshev coefficients in t
    (evaluated by Clenshaw's recurrence; monomial form loses 3 digits at this degree), 'rc2', 'nint', 'nlog'.
    """
    cheb = np.polynomial.chebyshev
    tn = np.cos(np.pi*(np.arange(ndeg) + 0.5)/ndeg)          # Chebyshev nodes in t
    S = np.zeros((nint, ndeg))
    A = np.zeros((nint, ndeg))
    rc2 = rc*rc
    dz = rc2/nint
    for i in range(nint):
        z = (i + 0.5*(tn + 1.0))*dz
        r = np.sqrt(z)
        if i < nlog:
            if kind == 0:
                x = xi*xi*z
                # Ein(x) = sum_{n>=1} (-1)^(n+1) x^n/(n n!)  (x <= 4.1 here)
                ein = np.zeros_like(x)
                term = np.ones_like(x)
                for n in range(1, 60):
                    term = term*(-x)/n
                    ein -= term/n
                s = (-np.log(xi*xi) - EULER + ein)/(4*np.pi)
                a = np.full_like(z, 1.0/(4*np.pi))
            else:
                a = ss.i0(kappa*r)/(4*np.pi)
                s = ss.k0(kappa*r)/(2*np.pi) - _gfar(kappa, xi, r) + a*np.log(z)
        else:
            s = _gloc(kind, kappa, xi, r)
            a = np.zeros_like(z)
        S[i] = cheb.chebfit(tn, s, ndeg - 1)
        A[i] = cheb.chebfit(tn, a, ndeg - 1)
    # derivative tables d/dz (z = r^2) of both series, for gradients: dG/dz = S' - A' log z - A/z and
    # grad G_loc = 2 (x - y) dG/dz.  d/dz = (2 nint / rc^2) d/dt on every interval.
    scale = 2.0*nint/rc2
    dS = np.zeros((nint, ndeg))
    dA = np.zeros((nint, ndeg))
    dS[:, :ndeg - 1] = np.array([cheb.chebder(S[i]) for i in range(nint)])*scale
    dA[:, :ndeg - 1] = np.array([cheb.chebder(A[i]) for i in range(nint)])*scale
    if kind == 0:
        # Laplace: fit the closed forms instead of differentiating the fits (which costs 2-3 digits):
        #   log-split intervals: A = 1/4pi (A' = 0), S' = xi^2 (1 - exp(-x))/(4 pi x), x = xi^2 z
        #   elsewhere:           dG/dz = -exp(-xi^2 z)/(4 pi z)
        for i in range(nint):
            z = (i + 0.5*(tn + 1.0))*dz
            x = xi*xi*z
            if i < nlog:
                val = xi*xi*(-np.expm1(-x))/(4*np.pi*x)
            else:
                val = -np.exp(-x)/(4*np.pi*z)
            dS[i] = cheb.chebfit(tn, val, ndeg - 1)
            dA[i] = 0.0
    tab = {'S': np.ascontiguousarray(S), 'A': np.ascontiguousarray(A), 'dS': np.ascontiguousarray(dS),
           'dA': np.ascontiguousarray(dA), 'rc2': rc2, 'nint': nint, 'nlog': nlog, 'ndeg': ndeg}
    if check:
        rng = np.random.default_rng(1)
        r = np.sqrt(rng.uniform(1e-6, 1.0, 400))*rc
        if kind == 0:
            ref = ss.exp1((xi*r)**2)/(4*np.pi)
        else:
            ref = np.where(r*r < nlog*dz, ss.k0(kappa*r)/(2*np.pi) - _gfar(kappa, xi, r), _gloc(kind, kappa, xi, r))
        err = np.abs(eval_local(tab, r*r) - ref).max()
        if not err < 2e-14*max(1.0, abs(np.log(xi))):
            raise Exception('Ewald local-kernel table inaccurate: %.2e (kappa*rc = %.2f)' % (err, kappa*rc))
    return tab


def eval_local(tab, z):
    """NumPy evaluation of the tables exactly as the device does it (z = r^2 < rc^2)."""
    z = np.asarray(z, dtype=float)
    u = np.clip(z*(tab['nint']/tab['rc2']), 0.0, tab['nint'] - 1e-9)
    i = np.minimum(u.astype(np.int64), tab['nint'] - 1)
    t = 2.0*(u - i) - 1.0
    def clenshaw(C):
        b0 = np.zeros_like(z); b1 = np.zeros_like(z)
        for c in range(tab['ndeg'] - 1, 0, -1):
            b0, b1 = 2.0*t*b0 - b1 + C[i, c], b0
        return t*b0 - b1 + C[i, 0]
    s = clenshaw(tab['S'])
    a = clenshaw(tab['A'])
    with np.errstate(divide='ignore', invalid='ignore'):
        lg = np.where(i < tab['nlog'], np.log(np.maximum(z, 1e-300)), 0.0)
    return s - a*lg


def eval_local_dz(tab, z):
    """d G_loc / d z from the derivative tables (NumPy statement of the device evaluation)"""
    z = np.asarray(z, dtype=float)
    u = np.clip(z*(tab['nint']/tab['rc2']), 0.0, tab['nint'] - 1e-9)
    i = np.minimum(u.astype(np.int64), tab['nint'] - 1)
    t = 2.0*(u - i) - 1.0
    def clenshaw(C):
        b0 = np.zeros_like(z); b1 = np.zeros_like(z)
        for c in range(tab['ndeg'] - 1, 0, -1):
            b0, b1 = 2.0*t*b0 - b1 + C[i, c], b0
        return t*b0 - b1 + C[i, 0]
    ds, da, a = clenshaw(tab['dS']), clenshaw(tab['dA']), clenshaw(tab['A'])
    zs = np.maximum(z, 1e-300)
    lg = np.where(i < tab['nlog'], np.log(zs), 0.0)
    return ds - da*lg - np.where(i < tab['nlog'], a/zs, 0.0)


def build_tiles(sx, sy, x0, y0, h, px, py, rc):
    """
    Tiles of TILE x TILE nodes of the (px, py) FFT box (node (i, j) at (x0 + i h, y0 + j h)) that have a source
    within rc of their cell span; returns (tile_ix, tile_iy, ptr, src, tile_map) with `src[ptr[t]:ptr[t+1]]`
    the sources within rc of tile t and tile_map (ntx, nty) int32 -> tile id or -1.
    """
    sx = np.asarray(sx, dtype=float); sy = np.asarray(sy, dtype=float)
    ntx, nty = -(-px//TILE), -(-py//TILE)
    span = TILE*h
    reach = int(np.ceil(rc/span)) + 1
    fx, fy = (sx - x0)/span, (sy - y0)/span
    cx, cy = np.floor(fx).astype(np.int64), np.floor(fy).astype(np.int64)
    offs = np.arange(-reach, reach + 1)
    OX, OY = np.meshgrid(offs, offs, indexing='ij')
    TX = (cx[:, None] + OX.ravel()[None, :])
    TY = (cy[:, None] + OY.ravel()[None, :])
    # distance from the source to the closed box [TX, TX+1] x [TY, TY+1] (tile units)
    ddx = np.maximum(0.0, np.maximum(TX - fx[:, None], fx[:, None] - (TX + 1)))
    ddy = np.maximum(0.0, np.maximum(TY - fy[:, None], fy[:, None] - (TY + 1)))
    ok = (ddx*ddx + ddy*ddy)*span*span <= rc*rc*(1.0 + 1e-12)
    ok &= (TX >= 0) & (TX < ntx) & (TY >= 0) & (TY < nty)
    sidx = np.broadcast_to(np.arange(sx.size)[:, None], TX.shape)[ok]
    key = (TX[ok]*nty + TY[ok])
    order = np.argsort(key, kind='stable')
    key, sidx = key[order], sidx[order]
    uniq, start = np.unique(key, return_index=True)
    ptr = np.concatenate([start, [key.size]]).astype(np.int32)
    tile_ix = (uniq//nty).astype(np.int32)
    tile_iy = (uniq % nty).astype(np.int32)
    tile_map = np.full(ntx*nty, -1, dtype=np.int32)
    tile_map[uniq] = np.arange(uniq.size, dtype=np.int32)
    return tile_ix, tile_iy, ptr, sidx.astype(np.int32), tile_map.reshape(ntx, nty)


def choose_box(lo, hi, h, m, x_origin, span, next_size):
    """
    FFT box along one axis.  [lo, hi]: extent of sources and targets; nodes are x_origin + i h (aligned with the
    background grid).  Returns (x0, p): first node and node count, with period p h >= padded extent + span.
    """
    i_lo = int(np.floor((lo - x_origin)/h)) - m - 1
    i_hi = int(np.ceil((hi - x_origin)/h)) + m + 1
    ext = i_hi - i_lo + 1
    need = ext + int(np.ceil(span/h)) + 1
    p = next_size(need)
    # the power-of-two FFT kernels move ~2.5x the bytes per second of the mixed-radix ones: prefer 2^a when the
    # extra area does not eat that up
    p2 = 1 << int(np.ceil(np.log2(need)))
    if p2 != p and float(p2)**2 < 2.5*float(p)**2:
        p = p2
    return x_origin + i_lo*h, p, i_lo
