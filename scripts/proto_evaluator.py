"""NumPy prototype of the free-space grid evaluator (parameter study; the CUDA version mirrors this)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, scipy.special as ss
from ipde_b200.slepian import SlepianMollifier
cheb = np.polynomial.chebyshev

def mollifier_series(r_sl, deg=120):
    S = SlepianMollifier(r_sl)
    x, _ = cheb.chebgauss(4*deg)
    c = cheb.chebfit(x, S.step(x), deg)
    return c, cheb.chebder(c), cheb.chebder(c, 2)

def trunc_sgf(kind, k, L, hk):
    if kind == 'laplace':
        out = np.empty_like(k)
        z = k == 0
        kn = np.where(z, 1.0, k)
        out = (1.0 - ss.j0(L*kn))/kn**2 - L*np.log(L)*ss.j1(L*kn)/kn
        out[z] = -L**2*np.log(L) + L**2*(1 + 2*np.log(L))/4   # reference formula (x 1/(2pi) convention folded below)
        return out
    return (1.0 + L*k*ss.j1(L*k)*ss.k0(L*hk) - L*hk*ss.j0(L*k)*ss.k1(L*hk))/(k**2 + hk**2)

def run(N=96, sw=20, r_sl=24, kind='mh', hk=2.0, seed=0):
    h = 2.4/N
    xv = -1.2 + h*np.arange(N)
    rng = np.random.default_rng(seed)
    ns = 150
    th = np.sort(rng.uniform(0, 2*np.pi, ns))
    sx, sy = 0.8*np.cos(th) + 0.013, 0.75*np.sin(th) - 0.007
    q = rng.standard_normal(ns)
    # targets: grid nodes inside radius 0.62, and off-grid points
    X, Y = np.meshgrid(xv, xv, indexing='ij')
    sel = (X/0.8)**2 + (Y/0.75)**2 < 0.80**2
    tx = np.concatenate([X[sel], rng.uniform(-0.4, 0.4, 40)]); ty = np.concatenate([Y[sel], rng.uniform(-0.4, 0.4, 40)])
    def G(r):
        return -np.log(r)/(2*np.pi) if kind == 'laplace' else ss.k0(hk*r)/(2*np.pi)
    def Gp(r):
        return -1.0/(2*np.pi*r) if kind == 'laplace' else -hk*ss.k1(hk*r)/(2*np.pi)
    d = np.hypot(tx[:, None] - sx, ty[:, None] - sy)
    exact = G(d).dot(q)
    print('min target-source distance / h: %.2f' % (d.min()/h))
    c0, c1, c2 = mollifier_series(r_sl)
    rc = sw*h
    def chi_all(r):
        z = 1 - 2*r/rc
        return cheb.chebval(z, c0), cheb.chebval(z, c1)*(-2/rc), cheb.chebval(z, c2)*(4/rc**2)
    def rho(r):
        r = np.maximum(r, 1e-300)
        ch, ch1, ch2 = chi_all(r)
        out = 2*ch1*Gp(r) + (ch2 + ch1/r)*G(r)
        return np.where((r < rc) & (r > 1e-12), out, 0.0)
    def gloc(r):
        ch, _, _ = chi_all(np.minimum(r, rc))
        return np.where(r < rc, ch*G(np.maximum(r, 1e-300)), 0.0)
    # padded grid
    P = N + 2*sw
    P2 = 2*P
    off = (P2 - N)//2
    xp = xv[0] + h*(np.arange(P2) - off)
    XP, YP = np.meshgrid(xp, xp, indexing='ij')
    op = np.zeros((P2, P2))
    for s in range(ns):
        dd = np.hypot(XP - sx[s], YP - sy[s])
        op += rho(dd)*q[s]
    # truncated kernel on 2*P2 grid
    B = 2*P2
    L = 2.5*N*h
    kv = np.fft.fftfreq(B, h/(2*np.pi))
    KX, KY = np.meshgrid(kv, kv, indexing='ij')
    ts = trunc_sgf(kind, np.hypot(KX, KY), L, hk)
    if kind == 'laplace':
        ts = ts/(1.0)   # reference laplace trunc_sgf is for G = -log r / (2 pi) with 2pi folded in
    Tbig = np.fft.ifft2(ts).real          # kernel centred at index 0, includes h^2
    idx = np.fft.fftfreq(P2, 1.0/P2).astype(int)   # 0..P2/2-1, -P2/2..-1
    T = Tbig[np.ix_(idx % B, idx % B)]
    TH = np.fft.fft2(T).real
    ufar = np.fft.ifft2(np.fft.fft2(op)*TH).real
    # evaluate at targets: grid nodes directly, off-grid by trigonometric interpolation of ufar (exact here)
    ngrid = sel.sum()
    ii, jj = np.nonzero(sel)
    far_grid = ufar[ii + off, jj + off]
    loc = gloc(d).dot(q)
    err_grid = np.abs(far_grid + loc[:ngrid] - exact[:ngrid]).max()/np.abs(exact).max()
    # spectrum decay of op
    oh = np.abs(np.fft.fft2(op))
    nyq = oh[P2//2 - 2:P2//2 + 2, :].max()/oh.max()
    # off-grid: NUDFT of ufar spectrum
    uh = np.fft.fft2(ufar)
    kk = np.fft.fftfreq(P2, 1.0/P2)
    ang_x = (tx[ngrid:] - xp[0])/(P2*h)*2*np.pi; ang_y = (ty[ngrid:] - xp[0])/(P2*h)*2*np.pi
    A = np.exp(1j*np.outer(ang_x, kk)); Bm = np.exp(1j*np.outer(ang_y, kk))
    far_off = np.einsum('jk,jk->j', A.dot(uh), Bm).real/P2**2
    err_off = np.abs(far_off + loc[ngrid:] - exact[ngrid:]).max()/np.abs(exact).max()
    # how band limited is ufar: fraction of spectrum above 0.75 nyquist
    m = (np.abs(kk)[:, None] > 0.75*P2/2) | (np.abs(kk)[None, :] > 0.75*P2/2)
    hi = np.abs(uh)[m].max()/np.abs(uh).max()
    print('%s N=%d sw=%d r_sl=%s: err grid %.2e  off-grid %.2e   op spectrum at nyquist %.1e, ufar spectrum >0.75nyq %.1e'
          % (kind, N, sw, r_sl, err_grid, err_off, nyq, hi))

if __name__ == '__main__':
    for kind in ['mh', 'laplace']:
        for sw, r_sl in [(20, 40), (20, 24), (20, 20), (24, 28), (24, 24), (28, 30)]:
            run(96, sw, r_sl, kind)
