"""
ORACLE tooling -- generates tests/golden/*.npz by running the REFERENCE's own code.

Run in the build container only (needs /root/reference; the GPU box does not have
it):      python oracle/make_golden.py

The importable part of the reference (ipde.annular.*, ipde.utilities,
ipde.derivatives, ipde.slepian.*) needs three shims for absent third-party
modules; they are injected into sys.modules here and
nothing of the reference is copied into the repository:
  * numexpr.evaluate                -> eval in the caller's frame
  * personal_utilities.scipy_gmres  -> right_gmres on scipy.sparse.linalg.gmres
  * np.Inf / np.product             -> NumPy-2 aliases
Everything written is small, seeded and deterministic.
"""
import os
import sys
import types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
REF = '/root/reference'


def install_shims():
    np.Inf = np.inf
    np.product = np.prod
    ne = types.ModuleType('numexpr')

    def evaluate(expr, local_dict=None, global_dict=None):
        f = sys._getframe(1)
        d = dict(f.f_globals)
        d.update(f.f_locals)
        return eval(expr, {'__builtins__': {}}, d)
    ne.evaluate = evaluate
    sys.modules['numexpr'] = ne
    pu = types.ModuleType('personal_utilities')
    sg = types.ModuleType('personal_utilities.scipy_gmres')
    import scipy.sparse.linalg as spla

    def right_gmres(A, b, M=None, tol=1e-8, restart=None, maxiter=None, verbose=False, **kw):
        n = b.shape[0]
        restart = 20 if restart is None else restart
        maxiter = 100 if maxiter is None else maxiter
        Mop = (lambda x: x) if M is None else M.matvec
        AM = spla.LinearOperator((n, n), dtype=complex, matvec=lambda x: A.matvec(Mop(x)))
        res = []
        y, info = spla.gmres(AM, b, rtol=tol, atol=0.0, restart=restart, maxiter=maxiter,
                             callback=lambda r: res.append(r), callback_type='pr_norm')
        return Mop(y), info, res
    sg.right_gmres = right_gmres
    sg.gmres = right_gmres
    pu.scipy_gmres = sg
    sys.modules['personal_utilities'] = pu
    sys.modules['personal_utilities.scipy_gmres'] = sg
    sys.path.insert(0, REF)


def star_geometry(n, a=0.1, f=5):
    t = np.linspace(0, 2*np.pi, n, endpoint=False)
    c = (1 + a*np.cos(f*t))*np.exp(1j*t)
    k = np.fft.fftfreq(n, 1.0/n)
    ch = np.fft.fft(c)
    cp = np.fft.ifft(ch*1j*k)
    cpp = np.fft.ifft(ch*(1j*k)**2)
    sp = np.abs(cp)
    nc = -1j*cp/sp
    cur = -(np.conj(cpp)*nc).real/sp**2
    return c, nc, sp, cur


def golden_annular_scalar():
    from ipde.annular.annular_full import ApproximateAnnularGeometry, RealAnnularGeometry
    from ipde.annular.modified_helmholtz import AnnularModifiedHelmholtzSolver
    from ipde.annular.poisson import AnnularPoissonSolver
    out = {}
    n, M, width = 96, 10, 0.15
    c, nc, sp, cur = star_geometry(n)
    AAG = ApproximateAnnularGeometry(n, M, width, np.mean(np.abs(c)))
    RAG = RealAnnularGeometry(sp, cur, AAG)
    CO = AAG.CO
    for name in ['D00', 'D01', 'D12', 'R01', 'R12', 'R02', 'P10', 'VI1', 'ibc_dirichlet',
                 'obc_dirichlet', 'ibc_neumann', 'obc_neumann']:
        out['CO_' + name] = getattr(CO, name)
    out['rv0'] = AAG.rv0
    out['psi1'] = RAG.psi1
    rx = c.real + AAG.rv0[:, None]*nc.real
    ry = c.imag + AAG.rv0[:, None]*nc.imag
    kk = 3.0
    sol = lambda x, y: np.exp(np.sin(kk*x))*np.sin(kk*y)
    lap = lambda x, y: kk**2*np.exp(np.sin(kk*x))*np.sin(kk*y)*(np.cos(kk*x)**2 - np.sin(kk*x) - 1.0)
    bx, by = c.real, c.imag
    ix, iy = c.real - width*nc.real, c.imag - width*nc.imag
    out.update(n=n, M=M, width=width, speed=sp, curvature=cur, approx_r=np.mean(np.abs(c)),
               radial_x=rx, radial_y=ry)
    # modified Helmholtz, k = 2.5, Dirichlet data from the exact solution
    hk = 2.5
    S = AnnularModifiedHelmholtzSolver(AAG, k=hk)
    f = hk**2*sol(rx, ry) - lap(rx, ry)
    u = S.solve(RAG, f, sol(ix, iy), sol(bx, by), tol=1e-14, maxiter=100, restart=50)
    out.update(mh_k=hk, mh_f=f, mh_ig=sol(ix, iy), mh_og=sol(bx, by), mh_u=u, mh_exact=sol(rx, ry),
               mh_iters=S.iterations_last_call)
    rng = np.random.default_rng(0)
    v = rng.standard_normal(M*n) + 1j*rng.standard_normal(M*n)
    out['mh_apply_in'] = v
    out['mh_apply_out'] = S._apply(v)
    out['mh_prec_out'] = np.array(S._optim_preconditioner(v))
    # zero-BC solve with random smooth forcing (what the helpers do)
    u0 = S.solve(RAG, f, np.zeros(n), np.zeros(n), tol=1e-14, maxiter=100, restart=50)
    out['mh_u_zero_bc'] = u0
    # Poisson
    P = AnnularPoissonSolver(AAG)
    fp = lap(rx, ry)
    up = P.solve(RAG, fp, sol(ix, iy), sol(bx, by), tol=1e-14, maxiter=100, restart=50)
    out.update(po_f=fp, po_u=up, po_iters=P.iterations_last_call)
    np.savez_compressed(os.path.join(GOLD, 'annular_scalar.npz'), **out)
    print('annular MH err vs exact', np.abs(u - sol(rx, ry)).max(), 'its', S.iterations_last_call)
    print('annular Poisson err vs exact', np.abs(up - sol(rx, ry)).max(), 'its', P.iterations_last_call)


def golden_annular_stokes():
    from ipde.annular.annular import ApproximateAnnularGeometry, RealAnnularGeometry
    from ipde.annular.stokes import AnnularStokesSolver
    out = {}
    n, M, width = 64, 8, 0.2
    c, nc, sp, cur = star_geometry(n, a=0.08, f=3)
    AAG = ApproximateAnnularGeometry(n, M, width, np.mean(np.abs(c)))
    RAG = RealAnnularGeometry(sp, cur, AAG)
    S = AnnularStokesSolver(AAG, mu=1.0)
    rng = np.random.default_rng(11)
    t = np.linspace(0, 2*np.pi, n, endpoint=False)
    rv = AAG.rv0[:, None]
    fr = np.cos(2*t)[None, :]*np.exp(rv) + 0.3*np.sin(t)[None, :]
    ft = np.sin(3*t)[None, :]*(1 + rv) - 0.2*np.cos(t)[None, :]*rv**2
    irg, itg = 0.1*np.cos(t), 0.05*np.sin(2*t)
    org, otg = 0.2*np.sin(t), -0.1*np.cos(3*t)
    ur, ut, p = S.solve(RAG, fr, ft, irg, itg, org, otg, tol=1e-14, restart=80, maxiter=4)
    # one operator / preconditioner application on a Hermitian-symmetric random vector
    def herm(rows):
        f = rng.standard_normal((rows, n))
        fh = np.fft.fft(f, axis=1)
        return np.concatenate([fh[:, :n//2], fh[:, n//2 + 1:]], axis=1)
    x = np.concatenate([herm(M).ravel(), herm(M).ravel(), herm(M - 1).ravel()])
    S.RAG = RAG
    S.combo1 = 2*RAG.DR_psi2*RAG.inv_psi2**2
    S.combo2 = RAG.DR_psi2**2*RAG.inv_psi2**2
    S.ik = AAG.k*1j
    S.iks = AAG.ks*1j
    out.update(n=n, M=M, width=width, radius=np.mean(np.abs(c)), speed=sp, curvature=cur, fr=fr, ft=ft,
               irg=irg, itg=itg, org=org, otg=otg, ur=ur, ut=ut, p=p, x=x,
               Ax=S._apply_optim_real(x), Px=S._preconditioner(x), Ax_complex=S._apply(x))
    np.savez_compressed(os.path.join(GOLD, 'annular_stokes.npz'), **out)
    print('annular_stokes.npz written')


def golden_slepian():
    from ipde.slepian.chebeval_bump_step import SlepianMollifier
    x = np.linspace(-1.25, 1.25, 501)
    out = {'x': x}
    for r in [4, 8, 16, 24, 30, 32, 40, 60]:
        S = SlepianMollifier(r)
        out['step_%d' % r] = S.step(x.copy())
        out['bump_%d' % r] = S.bump(x.copy())
    np.savez_compressed(os.path.join(GOLD, 'slepian_golden.npz'), **out)


def golden_misc():
    from ipde.derivatives import fd_x_4, fd_y_4, fourier
    from ipde.utilities import mfft, mifft, fourier_multiply
    rng = np.random.default_rng(1)
    f = rng.standard_normal((24, 20))
    h = 0.1
    kx = np.fft.fftfreq(24, h/(2*np.pi))[:, None]
    out = dict(f=f, h=h, fdx=fd_x_4(f, h), fdy=fd_y_4(f, h), fdx_p=fd_x_4(f, h, True),
               fdy_p=fd_y_4(f, h, True), fourier_x=fourier(f, 1j*kx))
    g = rng.standard_normal((5, 16))
    m = rng.standard_normal((5, 16))
    gh = mfft(g)
    out.update(g=g, m=m, mfft_g=gh, mifft_gh=mifft(gh), fm=fourier_multiply(gh, m))
    np.savez_compressed(os.path.join(GOLD, 'misc_golden.npz'), **out)


if __name__ == '__main__':
    install_shims()
    os.makedirs(GOLD, exist_ok=True)
    golden_annular_scalar()
    golden_annular_stokes()
    golden_slepian()
    golden_misc()
    print('golden vectors written to', GOLD)
