"""CPU: the oracle against golden vectors produced by the reference's own code (oracle/make_golden.py)."""
import os
import numpy as np
import pytest
from oracle.annular import (ApproximateAnnularGeometry, RealAnnularGeometry,
                            AnnularModifiedHelmholtzSolver, AnnularPoissonSolver)
from oracle import solvers as osol


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_chebyshev_operators_match_reference(golden_dir):
    g = _load(golden_dir, 'annular_scalar.npz')
    AAG = ApproximateAnnularGeometry(int(g['n']), int(g['M']), float(g['width']), float(g['approx_r']))
    for name in ['D00', 'D01', 'D12', 'R01', 'R12', 'R02', 'P10', 'VI1', 'ibc_dirichlet', 'obc_dirichlet',
                 'ibc_neumann', 'obc_neumann']:
        assert np.abs(getattr(AAG.CO, name) - g['CO_' + name]).max() < 1e-12, name
    assert np.abs(AAG.rv0 - g['rv0']).max() < 1e-15
    RAG = RealAnnularGeometry(g['speed'], g['curvature'], AAG)
    assert np.abs(RAG.psi1 - g['psi1']).max() < 1e-14


def test_annular_mh_apply_precond_solve_match_reference(golden_dir):
    g = _load(golden_dir, 'annular_scalar.npz')
    n = int(g['n'])
    AAG = ApproximateAnnularGeometry(n, int(g['M']), float(g['width']), float(g['approx_r']))
    RAG = RealAnnularGeometry(g['speed'], g['curvature'], AAG)
    S = AnnularModifiedHelmholtzSolver(AAG, float(g['mh_k']))
    S.RAG = RAG
    v = g['mh_apply_in']
    ref = g['mh_apply_out']
    assert np.abs(S._apply(v) - ref).max() <= 1e-13*np.abs(ref).max()
    refp = g['mh_prec_out']
    assert np.abs(S._preconditioner(v) - refp).max() <= 1e-11*np.abs(refp).max()
    u = S.solve(RAG, g['mh_f'], g['mh_ig'], g['mh_og'], tol=1e-14, maxiter=10, restart=50)
    assert np.abs(u - g['mh_u']).max() < 1e-13
    u0 = S.solve(RAG, g['mh_f'], np.zeros(n), np.zeros(n), tol=1e-14, maxiter=10, restart=50)
    assert np.abs(u0 - g['mh_u_zero_bc']).max() < 1e-13


def test_annular_poisson_matches_reference(golden_dir):
    g = _load(golden_dir, 'annular_scalar.npz')
    AAG = ApproximateAnnularGeometry(int(g['n']), int(g['M']), float(g['width']), float(g['approx_r']))
    RAG = RealAnnularGeometry(g['speed'], g['curvature'], AAG)
    P = AnnularPoissonSolver(AAG)
    up = P.solve(RAG, g['po_f'], g['mh_ig'], g['mh_og'], tol=1e-14, maxiter=10, restart=50)
    assert np.abs(up - g['po_u']).max() < 1e-13


def test_fd_derivatives_match_reference(golden_dir):
    g = _load(golden_dir, 'misc_golden.npz')
    assert np.abs(osol.fd_x_4(g['f'], float(g['h'])) - g['fdx']).max() < 1e-13
    assert np.abs(osol.fd_y_4(g['f'], float(g['h'])) - g['fdy']).max() < 1e-13


def test_nudft_is_inverse_dft_on_grid():
    rng = np.random.default_rng(0)
    f = rng.standard_normal((12, 10))
    fh = np.fft.fft2(f)
    x = 2*np.pi*np.arange(12)/12
    y = 2*np.pi*np.arange(10)/10
    X, Y = np.meshgrid(x, y, indexing='ij')
    v = osol.nudft2d_type2(fh, X.ravel(), Y.ravel())[0]
    assert np.abs(v.real - f.ravel()).max() < 1e-13


def test_radial_interpolation_reproduces_nodes():
    rng = np.random.default_rng(1)
    M, n = 8, 16
    # smooth function of (theta, t)
    th = (np.arange(M) + 0.5)*np.pi/M
    t = 2*np.pi*np.arange(n)/n
    fr = np.cos(2*th)[:, None]*np.sin(3*t)[None, :] + 0.3*np.cos(th)[:, None]
    T, TT = np.meshgrid(th, t, indexing='ij')
    v = osol.interpolate_radial_to_points(fr, T.ravel(), TT.ravel())
    assert np.abs(v - fr.ravel()).max() < 1e-13
    tq, tt = rng.uniform(0, np.pi, 20), rng.uniform(0, 2*np.pi, 20)
    v = osol.interpolate_radial_to_points(fr, tq, tt)
    assert np.abs(v - (np.cos(2*tq)*np.sin(3*tt) + 0.3*np.cos(tq))).max() < 1e-13


def test_oracle_annular_stokes_matches_reference_golden(golden_dir):
    """oracle/annular_stokes.py against inputs/outputs of the reference's own AnnularStokesSolver"""
    from oracle.annular import ApproximateAnnularGeometry, RealAnnularGeometry
    from oracle.annular_stokes import AnnularStokesSolver
    g = np.load(os.path.join(golden_dir, 'annular_stokes.npz'))
    AAG = ApproximateAnnularGeometry(int(g['n']), int(g['M']), float(g['width']), float(g['radius']), drop_nyquist=True)
    RAG = RealAnnularGeometry(g['speed'], g['curvature'], AAG)
    S = AnnularStokesSolver(AAG, 1.0)
    S.RAG = RAG
    rel = lambda a, b: np.abs(a - b).max()/np.abs(b).max()
    assert rel(S.apply(g['x']), g['Ax']) < 1e-13
    assert rel(S.preconditioner(g['x']), g['Px']) < 1e-11
    ur, ut, p = S.solve(RAG, g['fr'], g['ft'], g['irg'], g['itg'], g['org'], g['otg'], tol=1e-14, restart=80, maxiter=4)
    assert rel(ur, g['ur']) < 1e-12 and rel(ut, g['ut']) < 1e-12 and rel(p, g['p']) < 1e-11


def test_oracle_dirichlet_correction_reproduces_manufactured_solution():
    """SURVEY 8(f).1 on the CPU: oracle inhomogeneous solve + homogeneous Dirichlet correction lands on the analytic
    solution of examples/interior_modified_helmholtz.py:49-52 (what pins this step in the absence of pybie2d)"""
    from problems import single_star, manufactured
    from oracle.solvers import OracleScalarSolver
    from oracle.homogeneous import OracleDirichletCorrection
    from ipde_b200.qfs import Modified_Helmholtz_QFS
    from ipde_b200.embedded_function import EmbeddedFunction, BoundaryFunction
    k = 2.0
    ebdyc = single_star(nb=300, M=8)
    f, ua, sol = manufactured(ebdyc, k, kk=np.pi/3)
    ebdy = ebdyc[0]
    tolq = getattr(ebdy, 'qfs_tolerance', 1e-12)
    pair = (Modified_Helmholtz_QFS(ebdy.interface, True, True, True, k, tol=min(tolq, 1e-12), u2s=False, device_build=False),
            Modified_Helmholtz_QFS(ebdy.interface, False, True, True, k, tol=min(tolq, 1e-12), device_build=False))
    O = OracleScalarSolver(ebdyc, 'modified_helmholtz', k, 'spectral', [pair])
    uo = EmbeddedFunction(ebdyc, linear_data=np.asarray(O(f, tol=1e-13, restart=60, maxiter=5)))
    qb = Modified_Helmholtz_QFS(ebdy.bdy, True, True, False, k, tol=1e-12, b2c=False, device_build=False)
    corr = OracleDirichletCorrection(ebdyc, 'modified_helmholtz', k, [qb])
    u = corr(uo, BoundaryFunction(ebdyc, function=sol))
    err = np.abs(u - np.asarray(ua)).max()
    assert np.abs(np.asarray(uo) - np.asarray(ua)).max() > 0.1      # the correction matters
    assert err < 1e-6, err          # 2.2e-7 at this resolution (n_b = 300, M = 8)


def test_oracle_neumann_correction_reproduces_manufactured_solution():
    """the same with Neumann data (examples/multi_modified_helmholtz_neumann.py:97-163): oracle inhomogeneous solve +
    the correction whose normal derivative matches bc_n - d/dn ue lands on the analytic solution"""
    from problems import single_star, manufactured
    from oracle.solvers import OracleScalarSolver
    from oracle.homogeneous import OracleNeumannCorrection
    from ipde_b200.qfs import Modified_Helmholtz_QFS
    from ipde_b200.embedded_function import EmbeddedFunction
    k, kk = 2.0, np.pi/3
    ebdyc = single_star(nb=300, M=8)
    f, ua, sol = manufactured(ebdyc, k, kk=kk)
    ebdy = ebdyc[0]
    pair = (Modified_Helmholtz_QFS(ebdy.interface, True, True, True, k, tol=1e-12, u2s=False, device_build=False),
            Modified_Helmholtz_QFS(ebdy.interface, False, True, True, k, tol=1e-12, device_build=False))
    O = OracleScalarSolver(ebdyc, 'modified_helmholtz', k, 'spectral', [pair])
    uo = EmbeddedFunction(ebdyc, linear_data=np.asarray(O(f, tol=1e-13, restart=60, maxiter=5)))
    b = ebdy.bdy
    ux = kk*np.cos(kk*b.x)*np.exp(np.sin(kk*b.x))*np.sin(kk*b.y)
    uy = kk*np.exp(np.sin(kk*b.x))*np.cos(kk*b.y)
    bcn = ux*b.normal_x + uy*b.normal_y
    qb = Modified_Helmholtz_QFS(b, True, True, False, k, tol=1e-12, u2s=False, device_build=False)
    corr = OracleNeumannCorrection(ebdyc, k, [qb])
    u = corr(uo, bcn)
    err = np.abs(u - np.asarray(ua)).max()
    assert np.abs(np.asarray(uo) - np.asarray(ua)).max() > 0.1      # the correction matters
    assert 1 <= corr.iterations <= 40
    assert err < 2e-6, err


def test_oracle_stokes_reproduces_published_errors():
    """examples/interior_stokes_with_multi.py at its published resolution nb = 300 (errs_U 6.65e-07, errs_p 6.57e-05,
    :31-32) on the CPU: oracle Stokes solve + the homogeneous velocity correction (:98-118) + pressure normalisation
    (:120-125).  The pressure figure needs the continuity step of OracleVectorHelper.correct: without it the annulus
    is off by 7.5e-02."""
    from problems import stokes_manufactured
    from ipde_b200.geometry import GSB, Grid, star, arc_length_parameterize
    from ipde_b200.embedded_boundary import EmbeddedBoundary
    from ipde_b200.ebdy_collection import EmbeddedBoundaryCollection
    from ipde_b200.embedded_function import EmbeddedFunction
    from ipde_b200.slepian import SlepianMollifier
    from ipde_b200.qfs import Stokes_QFS
    from oracle.solvers import OracleVectorSolver, layer_apply_stokes
    nb, M = 300, 12
    bdy = GSB(c=star(nb, a=0.1, f=5))
    bdy = GSB(*arc_length_parameterize(bdy.x, bdy.y))
    Lb = np.pi/2
    grid = Grid([-Lb, Lb], nb//2, [-Lb, Lb], nb//2)
    MOL = SlepianMollifier(int(1.5*M))
    ebdy = EmbeddedBoundary(bdy, True, M, grid.xh*0.75, heaviside=MOL.step)
    ebdyc = EmbeddedBoundaryCollection([ebdy])
    ebdyc.register_grid(grid)
    ebdyc.ready_bump(MOL.bump, (Lb - ebdy.radial_width, Lb - ebdy.radial_width), ebdy.radial_width)
    fu, fv, ua, va, pa, (u_f, v_f, p_f) = stokes_manufactured(ebdyc)
    pair = (Stokes_QFS(ebdy.interface, True, True, True, tol=1e-13, u2s=False, device_build=False),
            Stokes_QFS(ebdy.interface, False, True, True, tol=1e-13, device_build=False))
    O = OracleVectorSolver(ebdyc, 'spectral', [pair])
    E = lambda a: EmbeddedFunction(ebdyc, linear_data=np.asarray(a))
    u, v, p = [E(a) for a in O(fu, fv, tol=1e-14, restart=50, maxiter=4)]
    h = O.helpers[0]
    qb = Stokes_QFS(bdy, True, True, True, tol=1e-13, device_build=False)
    g = np.concatenate([u_f(bdy.x, bdy.y) - h._bv.dot(u[0]), v_f(bdy.x, bdy.y) - h._bv.dot(v[0])])
    sig = qb.u2s(g).reshape(2, -1)*qb.source.weights
    pts = ebdyc.grid_and_radial_pts
    cu, cv, cp = layer_apply_stokes(qb.source.x, qb.source.y, sig[0], sig[1], pts.x, pts.y)
    area = EmbeddedFunction(ebdyc, function=lambda x, y: np.ones_like(x)).integrate()
    pt = E(np.asarray(p) + cp)
    eP = np.abs((np.asarray(pt) - pt.integrate()/area) - (np.asarray(pa) - pa.integrate()/area)).max()
    eU = max(np.abs(np.asarray(u) + cu - np.asarray(ua)).max(), np.abs(np.asarray(v) + cv - np.asarray(va)).max())
    assert eU < 1.5*6.65e-07 and eP < 1.5*6.57e-05, (eU, eP)          # measured 6.61e-07 / 6.48e-05


@pytest.mark.parametrize('kind,kappa,N', [(0, 0.0, 96), (1, 2.0, 96), (1, 20.0, 64)])
def test_numpy_spectral_ewald_matches_direct_sum(kind, kappa, N):
    """host tables and mathematics of the free-space evaluator (ipde_b200/ewald.py) on the CPU: truncated kernel,
    Gaussian split, local Chebyshev tables -- against the all-pairs sum of the reference's kernels"""
    from oracle.ewald import ewald_against_direct_sum
    err_grid, err_off = ewald_against_direct_sum(N, kind, kappa)
    bound = 1e-12 if kappa >= 20.0 else 1e-13
    assert err_grid < bound and err_off < bound, (err_grid, err_off)


def test_oracle_dirichlet_correction_poisson():
    """the same for the Poisson path (examples/poisson_for_paper.py:88-105): log sources outside the domain"""
    from problems import single_star, manufactured
    from oracle.solvers import OracleScalarSolver
    from oracle.homogeneous import OracleDirichletCorrection
    from ipde_b200.qfs import Laplace_QFS
    from ipde_b200.embedded_function import EmbeddedFunction, BoundaryFunction
    ebdyc = single_star(nb=300, M=8, bump=True)
    f, ua, sol = manufactured(ebdyc, 0.0, kk=np.pi/3)
    e = ebdyc[0]
    pair = (Laplace_QFS(e.interface, True, True, True, tol=1e-12, u2s=False, device_build=False),
            Laplace_QFS(e.interface, False, True, True, tol=1e-12, device_build=False))
    O = OracleScalarSolver(ebdyc, 'poisson', 0.0, 'spectral', [pair])
    uo = EmbeddedFunction(ebdyc, linear_data=np.asarray(O(f, tol=1e-13, restart=60, maxiter=5)))
    qb = Laplace_QFS(e.bdy, True, True, False, tol=1e-12, b2c=False, device_build=False)
    u = OracleDirichletCorrection(ebdyc, 'poisson', 0.0, [qb])(uo, BoundaryFunction(ebdyc, function=sol))
    assert np.abs(u - np.asarray(ua)).max() < 2e-7        # 2.6e-8 at this resolution


def test_fast_oracle_matches_direct_oracle():
    """oracle/fast.py (FFT-based NUFFT + spectral-Ewald sums, the CPU baseline of bench.py) against the direct
    statement (NUDFT + all-pairs sums) on the same inputs"""
    from problems import three_bodies, manufactured
    from oracle.solvers import OracleScalarSolver, nudft2d_type2
    from oracle.fast import FastScalarSolver, nufft2d2
    from ipde_b200.qfs import Modified_Helmholtz_QFS
    rng = np.random.default_rng(0)
    fh = rng.standard_normal((2, 48, 40)) + 1j*rng.standard_normal((2, 48, 40))
    x, y = rng.uniform(0, 2*np.pi, 50), rng.uniform(0, 2*np.pi, 50)
    ref = nudft2d_type2(fh, x, y)
    assert np.abs(nufft2d2(fh, x, y) - ref).max() < 1e-13*np.abs(ref).max()
    ebdyc = three_bodies(nb=80, M=8)
    k = 5.0
    f, ua, _ = manufactured(ebdyc, k, kk=10*np.pi/7)
    pairs = [(Modified_Helmholtz_QFS(e.interface, e.interior, True, True, k, tol=1e-12, u2s=False, device_build=False),
              Modified_Helmholtz_QFS(e.interface, not e.interior, True, True, k, tol=1e-12, device_build=False))
             for e in ebdyc]
    gm = dict(tol=1e-13, restart=60, maxiter=5)
    uo = OracleScalarSolver(ebdyc, 'modified_helmholtz', k, 'spectral', pairs)(f, **gm)
    uf = FastScalarSolver(ebdyc, 'modified_helmholtz', k, pairs, workers=2)(f, **gm)
    assert np.abs(uf - uo).max() < 1e-12*np.abs(uo).max()


def test_interpolate_to_points_numpy_statement():
    """ebdyc.register_points + the reference's interpolation formulas (ipde/ebdy_collection.py:666-707) in NumPy: NUDFT
    of the smoothed grid part in zone 1, Chebyshev x Fourier interpolation in the annuli -- pins the partition, the
    affine maps and the accuracy the device path (tests/test_gpu_solvers.py::test_interpolate_to_points) is held to"""
    from problems import three_bodies
    from ipde_b200.embedded_function import EmbeddedFunction
    from ipde_b200.geometry import fourier_resample
    from oracle.solvers import nudft2d_type2, interpolate_radial_to_points
    ebdyc = three_bodies(nb=200, M=12)
    kk = 4*np.pi/7
    sol = lambda x, y: np.exp(np.sin(kk*x))*np.sin(kk*y)
    ff = EmbeddedFunction(ebdyc, function=sol)
    g = ebdyc.grid
    rng = np.random.default_rng(9)
    n = 1500
    x = rng.uniform(g.x_bounds[0], g.x_bounds[1], n)
    y = rng.uniform(g.y_bounds[0], g.y_bounds[1], n)
    p = ebdyc.registered_partitions[ebdyc.register_points(x, y)]
    out = np.full(n, np.nan)
    out[p.zone1] = nudft2d_type2(np.fft.fft2(ff.get_smoothed_grid_value())[None], p.x_transf, p.y_transf)[0].real
    for ind in range(len(ebdyc)):
        out[p.zone2l[ind]] = interpolate_radial_to_points(ff[ind], p.zone2_transfr[ind], p.zone2t[ind])
    ex = sol(x, y)
    z = x + 1j*y
    phys = np.ones(n, dtype=bool)
    for e in ebdyc:
        fine = fourier_resample(e.bdy.c, 8*e.bdy.N)
        ang = np.angle((np.roll(fine, -1)[None, :] - z[:, None])/(fine[None, :] - z[:, None])).sum(axis=1)/(2*np.pi)
        phys &= (np.abs(ang) > 0.5) if e.interior else (np.abs(ang) <= 0.5)
    z1, z2 = p.zone1 & phys, np.concatenate(p.zone2l)
    assert z1.sum() > 300 and z2.size > 100
    assert np.abs(out[z1] - ex[z1]).max() < 2e-8           # roll-off of the cutoff on the 480^2 grid: 4e-09
    assert np.abs(out[z2] - ex[z2]).max() < 1e-12


def test_cubic_interpolation_statement_properties():
    """oracle statement of fast_interp.interp2d(k=3, periodic) (absent package; ipde/ebdy_collection.py:594-603): the
    centred four-node Lagrange stencil is the only placement that (i) returns nodal values at the nodes, (ii) is exact
    for tensor cubics wherever the stencil does not wrap, (iii) is symmetric about the cell (mirrored data gives mirrored
    values) and (iv) converges at fourth order on a periodic function"""
    from oracle.solvers import lagrange_interp2d_periodic
    rng = np.random.default_rng(4)
    N = 40
    h = 2.0/N
    xv = -1.0 + h*np.arange(N)
    f = rng.standard_normal((N, N))
    ii, jj = rng.integers(0, N, 50), rng.integers(0, N, 50)
    v = lagrange_interp2d_periodic(f, -1.0, -1.0, h, h, xv[ii], xv[jj], 3)
    assert np.abs(v - f[ii, jj]).max() < 1e-13
    poly = lambda x, y: (1 + x - 0.5*x**2 + 2*x**3)*(2 - y + y**2 - 0.3*y**3)
    P = poly(xv[:, None], xv[None, :])
    px, py = rng.uniform(-0.8, 0.7, 200), rng.uniform(-0.8, 0.7, 200)
    assert np.abs(lagrange_interp2d_periodic(P, -1.0, -1.0, h, h, px, py, 3) - poly(px, py)).max() < 1e-12
    # mirror symmetry: x -> -x about a cell centre maps the stencil onto itself
    px2 = rng.uniform(-0.5, 0.5, 100)
    a = lagrange_interp2d_periodic(f, -1.0, -1.0, h, h, px2, py[:100], 3)
    fm = f[::-1]                                  # node i <-> node N-1-i : x -> -x - h + ... (grid reflected)
    xm = (xv[0] + xv[-1]) - px2
    b = lagrange_interp2d_periodic(fm, -1.0, -1.0, h, h, xm, py[:100], 3)
    assert np.abs(a - b).max() < 1e-12
    errs = []
    for n in (32, 64, 128):
        hh = 2.0/n
        g = -1.0 + hh*np.arange(n)
        F = np.exp(np.sin(np.pi*g))[:, None]*np.cos(np.pi*g)[None, :]
        q = lagrange_interp2d_periodic(F, -1.0, -1.0, hh, hh, px, py, 3)
        errs.append(np.abs(q - np.exp(np.sin(np.pi*px))*np.cos(np.pi*py)).max())
    rate = np.log2(errs[0]/errs[1]), np.log2(errs[1]/errs[2])
    assert 3.7 < rate[0] < 4.4 and 3.7 < rate[1] < 4.4, (errs, rate)
