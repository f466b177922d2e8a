"""NumPy prototype of the Gaussian-split (Ewald) free-space grid evaluator."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, scipy.special as ss
cheb = np.polynomial.chebyshev

def gfar_table(kind, hk, xi, rc, deg=80):
    """Chebyshev coefficients in u = 2 r^2/rc^2 - 1 of the smooth far kernel G_far(r)."""
    t, _ = cheb.chebgauss(2*deg)
    r = rc*np.sqrt((t + 1)/2)
    if kind == 'laplace':
        val = -(np.log(r*r) + ss.exp1(xi*xi*r*r))/(4*np.pi)
    else:
        # (1/4pi) int_0^inf exp(-r^2 xi^2 e^-s - (k^2/4xi^2) e^s) ds, Gauss-Legendre on [0, smax]
        a = hk*hk/(4*xi*xi)
        smax = np.log(40.0/a) if a < 40 else 1.0
        x, w = np.polynomial.legendre.leggauss(400)
        s = 0.5*smax*(x + 1); w = 0.5*smax*w
        val = np.array([(np.exp(-(ri*xi)**2*np.exp(-s) - a*np.exp(s))*w).sum() for ri in r])/(4*np.pi)
    return cheb.chebfit(t, val, deg)

def trunc_sgf(kind, k, L, hk):
    if kind == 'laplace':
        z = k == 0
        kn = np.where(z, 1.0, k)
        out = (1.0 - ss.j0(L*kn))/kn**2 - L*np.log(L)*ss.j1(L*kn)/kn
        out[z] = L*L/4 - L*L*np.log(L)/2
        return out
    return (1.0 + L*k*ss.j1(L*k)*ss.k0(L*hk) - L*hk*ss.j0(L*k)*ss.k1(L*hk))/(k**2 + hk**2)

def run(N=96, sw=21, kind='mh', hk=2.0, eps=1e-14, seed=0):
    h = 2.4/N
    xv = -1.2 + h*np.arange(N)
    rng = np.random.default_rng(seed)
    ns = 150
    th = np.sort(rng.uniform(0, 2*np.pi, ns))
    sx, sy = 0.8*np.cos(th) + 0.013, 0.75*np.sin(th) - 0.007
    q = rng.standard_normal(ns)
    X, Y = np.meshgrid(xv, xv, indexing='ij')
    sel = (X/0.8)**2 + (Y/0.75)**2 < 0.80**2
    tx, ty = X[sel], Y[sel]
    G = (lambda r: -np.log(r)/(2*np.pi)) if kind == 'laplace' else (lambda r: ss.k0(hk*r)/(2*np.pi))
    d = np.hypot(tx[:, None] - sx, ty[:, None] - sy)
    exact = G(d).dot(q)
    rc = sw*h
    xi = np.sqrt(np.log(1/eps))/rc
    print('xi*h = %.3f (needs <= %.3f)' % (xi*h, np.pi/(2*np.sqrt(np.log(1/eps)))))
    c = gfar_table(kind, hk, xi, rc)
    print('cheb tail', np.abs(c[-4:]).max())
    gloc = lambda r: np.where(r < rc, G(r) - cheb.chebval(2*r*r/rc**2 - 1, c), 0.0)
    ck = np.exp(-hk*hk/(4*xi*xi)) if kind != 'laplace' else 1.0
    P2 = 2*(N + 2*sw)
    off = (P2 - N)//2
    xp = xv[0] + h*(np.arange(P2) - off)
    XP, YP = np.meshgrid(xp, xp, indexing='ij')
    op = np.zeros((P2, P2))
    for s in range(ns):
        dd2 = (XP - sx[s])**2 + (YP - sy[s])**2
        op += np.where(dd2 < rc*rc, q[s]*ck*xi*xi/np.pi*np.exp(-xi*xi*dd2), 0.0)
    B = 2*P2
    L = 2.5*N*h
    kv = np.fft.fftfreq(B, h/(2*np.pi))
    KX, KY = np.meshgrid(kv, kv, indexing='ij')
    Tbig = np.fft.ifft2(trunc_sgf(kind, np.hypot(KX, KY), L, hk)).real
    idx = np.fft.fftfreq(P2, 1.0/P2).astype(int)
    T = Tbig[np.ix_(idx % B, idx % B)]
    ufar = np.fft.ifft2(np.fft.fft2(op)*np.fft.fft2(T).real).real
    ii, jj = np.nonzero(sel)
    val = ufar[ii + off, jj + off] + gloc(d).dot(q)
    print('%s N=%d sw=%d: rel err %.2e' % (kind, N, sw, np.abs(val - exact).max()/np.abs(exact).max()))

if __name__ == '__main__':
    for kind in ['laplace', 'mh']:
        for sw, eps in [(16, 1e-10), (21, 1e-14), (22, 1e-15), (24, 1e-15)]:
            run(96, sw, kind, eps=eps)
    run(128, 21, 'mh', hk=5.0)
    run(128, 21, 'mh', hk=0.3)
