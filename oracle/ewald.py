"""
ORACLE (test infrastructure, CPU, NumPy) -- never imported by the product.

NumPy statement of the spectral-Ewald free-space evaluator that csrc/ewald.cu implements (mathematics and host
tables: ipde_b200/ewald.py): spread with the Gaussian w, multiply the FFT by the truncated kernel sampled directly on
the FFT grid (period >= extent + L), on-grid values via the extra factor exp(-k^2/8 xi^2), off-grid values by
gathering H with the same window, local part G - G_far within rc from the piecewise Chebyshev tables.  It replaces the
reference's Grid_Evaluator / Layer_Apply sums (ipde/grid_evaluators/scalar_grid_evaluator.py:188-231,299-307;
kernels laplace_grid_evaluator.py:8-12, modified_helmholtz_grid_evaluator.py:8-9) and is pinned by the all-pairs sum
of exactly those kernels (tests/test_oracle.py).
"""
import numpy as np
import scipy.special as ss

from ipde_b200.ewald import EwaldParameters, trunc_kernel_hat, local_tables, eval_local


def ewald_against_direct_sum(N=96, kind=1, hk=2.0, seed=0):
    """returns (relative error at grid targets, at off-grid targets) of the NumPy spectral-Ewald evaluation against
    the all-pairs sum, for 150 sources on an ellipse and targets inside it"""
    h = 2.4/N
    xv = -1.2 + h*np.arange(N)
    rng = np.random.default_rng(seed)
    ns = 150
    th = np.sort(rng.uniform(0, 2*np.pi, ns))
    sx, sy = 0.8*np.cos(th) + 0.013, 0.75*np.sin(th) - 0.007
    q = rng.standard_normal(ns)
    X, Y = np.meshgrid(xv, xv, indexing='ij')
    sel = (X/0.8)**2 + (Y/0.75)**2 < 0.80**2
    ngrid = sel.sum()
    tx = np.concatenate([X[sel], 0.62*np.cos(th[::3]) + 0.001]); ty = np.concatenate([Y[sel], 0.6*np.sin(th[::3])])
    G = (lambda r: -np.log(r)/(2*np.pi)) if kind == 0 else (lambda r: ss.k0(hk*r)/(2*np.pi))
    d = np.hypot(tx[:, None] - sx, ty[:, None] - sy)
    exact = G(d).dot(q)
    P = EwaldParameters(h)
    xi, rc, rw, m = P.xi, P.rc, P.rw, P.m
    D = 2*0.82 + 2*np.sqrt(2)*m*h
    L = D + 2*h
    P2 = int(np.ceil((N + 2*m)*1.0 + L/h)) + 1
    P2 += P2 % 2
    x0 = xv[0] - m*h
    xp = x0 + h*np.arange(P2)
    XP, YP = np.meshgrid(xp, xp, indexing='ij')
    ck = np.exp(-hk*hk/(4*xi*xi)) if kind == 1 else 1.0
    Sw = np.zeros((P2, P2))
    for s in range(ns):
        dd2 = (XP - sx[s])**2 + (YP - sy[s])**2
        Sw += np.where(dd2 < rw*rw, q[s]*ck*2*xi*xi/np.pi*np.exp(-2*xi*xi*dd2), 0.0)
    kv = np.fft.fftfreq(P2, h/(2*np.pi))
    KX, KY = np.meshgrid(kv, kv, indexing='ij')
    K2 = KX**2 + KY**2
    GT = trunc_kernel_hat(kind, hk, L, np.sqrt(K2))
    Sh = np.fft.fft2(Sw)
    U = np.fft.ifft2(GT*np.exp(-K2/(8*xi*xi))*Sh).real
    H = np.fft.ifft2(GT*Sh).real
    tab = local_tables(kind, hk, xi, rc)
    loc = (eval_local(tab, d**2)*(d < rc)).dot(q)
    ii, jj = np.nonzero(sel)
    val_grid = U[ii + m, jj + m] + loc[:ngrid]
    err_grid = np.abs(val_grid - exact[:ngrid]).max()/np.abs(exact).max()
    far_off = np.empty(tx.size - ngrid)
    for t in range(ngrid, tx.size):
        dd2 = (XP - tx[t])**2 + (YP - ty[t])**2
        far_off[t - ngrid] = h*h*(np.where(dd2 < rw*rw, 2*xi*xi/np.pi*np.exp(-2*xi*xi*dd2), 0.0)*H).sum()
    err_off = np.abs(far_off + loc[ngrid:] - exact[ngrid:]).max()/np.abs(exact).max()
    return err_grid, err_off

