"""GPU: every kernel of libipde_b200.so against NumPy/SciPy or the oracle (through the C ABI)."""
import os
import numpy as np
import pytest
import scipy.linalg
import scipy.sparse
import scipy.special

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max()/max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope='module')
def D():
    import torch
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    import ipde_b200.device as dev
    return dev


@pytest.mark.parametrize('n', [8, 60, 64, 79, 96, 100, 256, 474, 512, 1000, 1022, 1024, 1430, 2048, 4096, 7168, 8192,
                               30030, 768, 1536, 3072, 6144, 12288, 24576, 49152, 16384, 32768, 65536, 98304])
def test_fft_c2c(D, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((3, n)) + 1j*rng.standard_normal((3, n))
    p = D.FFT1D(n)
    xd = D.to_dev(x)
    y = p.c2c(xd, -1).cpu().numpy()
    assert rel(y, np.fft.fft(x)) < 1e-14*np.log2(n)
    z = p.c2c(D.to_dev(y), +1, 1.0/n).cpu().numpy()
    assert rel(z, x) < 1e-14*np.log2(n)
    # in place
    p.c2c(xd, -1, out=xd)
    assert rel(xd.cpu().numpy(), np.fft.fft(x)) < 1e-14*np.log2(n)


@pytest.mark.parametrize('n', [96, 256, 474, 512, 1024, 2048, 4096, 8192, 9000, 768, 3072, 6144, 12288, 16384, 24576,
                               49152])
def test_fft_mul_fft(D, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((5, n)) + 1j*rng.standard_normal((5, n))
    m = rng.standard_normal((5, n))
    p = D.FFT1D(n)
    y = p.mul_fft(D.to_dev(x), D.to_dev(m)).cpu().numpy()
    assert rel(y, np.fft.fft(m*np.fft.ifft(x))) < 1e-13


@pytest.mark.parametrize('shape', [(64, 64), (96, 80), (100, 36), (1024, 40), (2048, 128), (128, 2048), (4096, 64), (4096, 100),
                                   (6000, 20), (8192, 40), (16384, 36),
                                   (9, 512), (15, 1024), (20, 2048), (10, 4096), (6, 8192), (8, 16384),
                                   (474, 474), (94, 158), (2102, 34), (32768, 20), (7, 32768), (2500, 32768)])
def test_grid_fft2(D, shape):
    nx, ny = shape
    rng = np.random.default_rng(nx + ny)
    f = rng.standard_normal(shape)
    g = D.GridPlan(nx, ny, 0.1, 0.1)
    spec = g.fft2(D.to_dev(f))
    full = g.spec_to_numpy_full(spec)
    assert rel(full, np.fft.fft2(f)) < 3e-14
    back = g.ifft2(spec).cpu().numpy()
    assert rel(back, f) < 3e-14


@pytest.mark.parametrize('kind,k', [(1, 2.5), (0, 0.0)])
@pytest.mark.parametrize('shape', [(96, 80), (1024, 48), (2048, 64), (4096, 36), (8192, 36), (16384, 40), (474, 94),
                                   (32768, 36), (40, 32768), (24, 16384)])
def test_grid_solve_scalar(D, kind, k, shape):
    from ipde_b200._lib import lib
    nx, ny = shape
    xh, yh = 2.4/nx, 2.0/ny
    rng = np.random.default_rng(7)
    f = rng.standard_normal(shape)
    bumpy = np.abs(rng.standard_normal(shape))
    bumpy /= bumpy.sum()*xh*yh
    g = D.GridPlan(nx, ny, xh, yh)
    fc, uch, work, uc, bd = D.to_dev(f), g.new_spec(), g.new_spec(), g.new_real(), D.to_dev(bumpy)
    lib().ipde_grid_solve_scalar(g.h, kind, k, D.ptr(fc), D.ptr(bd), D.ptr(uch), D.ptr(work), D.ptr(uc), D.stream())
    kx = np.fft.fftfreq(nx, xh/(2*np.pi))[:, None]
    ky = np.fft.fftfreq(ny, yh/(2*np.pi))
    lap = -kx*kx - ky*ky
    if kind == 0:
        ff = f - f.sum()*xh*yh*bumpy
        lap[0, 0] = np.inf
        ref_h = np.fft.fft2(ff)/lap
    else:
        ref_h = np.fft.fft2(f)/(k*k - lap)
    ref = np.fft.ifft2(ref_h).real
    assert rel(uc.cpu().numpy(), ref) < 1e-12
    assert rel(g.spec_to_numpy_full(uch), ref_h) < 1e-12


@pytest.mark.parametrize('shape', [(48, 40), (2048, 32), (32, 1024), (40, 8192), (24, 16384)])
def test_nufft2_against_direct_sum(D, shape):
    # row lengths 2048 ... 32768 of the fine grid: the row pass stores only the 64-real blocks under the targets'
    # stencils (register kernel and the cluster-pair long-row kernel)
    from oracle.solvers import nudft2d_type2
    nx, ny = shape
    xh, yh = 2.4/nx, 3.0/ny
    rng = np.random.default_rng(3)
    f = rng.standard_normal(shape)
    g = D.GridPlan(nx, ny, xh, yh)
    spec = g.fft2(D.to_dev(f))
    x = rng.uniform(0, 2*np.pi, 37)
    y = rng.uniform(0, 2*np.pi, 37)
    x[0], y[0] = 0.0, 0.0
    x[1], y[1] = 2*np.pi*(1 - 1e-12), 1e-9
    nu = D.Nufft2(g, x, y)
    fh = np.fft.fft2(f)
    kx = np.fft.fftfreq(nx, xh/(2*np.pi))[:, None]
    ky = np.fft.fftfreq(ny, yh/(2*np.pi))
    ref = nudft2d_type2(np.stack([fh, 1j*kx*fh, 1j*ky*fh]), x, y).real
    scale = [np.abs(f).max(), np.abs(kx).max()*np.abs(f).max(), np.abs(ky).max()*np.abs(f).max()]
    tol = 2e-13 if max(shape) <= 2048 else 2e-12      # round-off of 10^5 ... 4 10^5 full-band random modes per point
    for i, (dx, dy) in enumerate([(0, 0), (1, 0), (0, 1)]):
        v = nu(spec, dx, dy).cpu().numpy()
        assert np.abs(v - ref[i]).max()/scale[i] < tol, (dx, dy)
    # value and gradient from ONE fine grid (gradient by differentiating the interpolation kernel); target 1 sits
    # 1e-12 from a fine-grid node, where the kernel derivative is singular at the edge of the stencil
    for i, v in enumerate(nu.apply3(spec)):
        assert np.abs(v.cpu().numpy() - ref[i]).max()/scale[i] < (tol if i == 0 else 5*tol), i


@pytest.mark.parametrize('kind,k', [(0, 0.0), (1, 3.7)])
def test_layer_apply_scalar_normal_derivative(D, kind, k):
    """k_layer_scalar_dn: d/dn_t of the point-source sums against NumPy/SciPy (K1), near and far pairs"""
    rng = np.random.default_rng(11)
    ns, nt = 700, 333
    sx, sy = rng.uniform(-1, 1, ns), rng.uniform(-1, 1, ns)
    tx, ty = rng.uniform(-1.2, 1.2, nt), rng.uniform(-1.2, 1.2, nt)
    tx[:5], ty[:5] = sx[:5] + 1e-3, sy[:5] - 2e-3          # close pairs (small-argument branch of K1)
    tx[5:8], ty[5:8] = sx[5:8] + 3.0, sy[5:8] + 2.0         # large arguments (Chebyshev tail)
    th = rng.uniform(0, 2*np.pi, nt)
    nx, ny = np.cos(th), np.sin(th)
    q = rng.standard_normal(ns)
    dx, dy = tx[:, None] - sx[None, :], ty[:, None] - sy[None, :]
    r = np.hypot(dx, dy)
    dn = dx*nx[:, None] + dy*ny[:, None]
    if kind == 0:
        ref = (-dn/r**2/(2*np.pi)).dot(q)
    else:
        ref = (-k*scipy.special.k1(k*r)*dn/r/(2*np.pi)).dot(q)
    out = D.layer_apply_scalar_dn(kind, k, D.to_dev(sx), D.to_dev(sy), D.to_dev(q), D.to_dev(tx), D.to_dev(ty),
                                  D.to_dev(nx), D.to_dev(ny)).cpu().numpy()
    assert np.abs(out - ref).max() < 1e-13*np.abs(ref).max()
    acc = D.to_dev(np.ones(nt))
    D.layer_apply_scalar_dn(kind, k, D.to_dev(sx), D.to_dev(sy), D.to_dev(q), D.to_dev(tx), D.to_dev(ty), D.to_dev(nx),
                            D.to_dev(ny), out=acc, accumulate=True)
    assert np.abs(acc.cpu().numpy() - 1.0 - ref).max() < 1e-13*np.abs(ref).max()


@pytest.mark.parametrize('M,n', [(8, 40), (16, 300), (20, 1000)])
def test_radial_to_points(D, M, n):
    from oracle.solvers import interpolate_radial_to_points
    rng = np.random.default_rng(M*n)
    fr = rng.standard_normal((M, n))
    th, t = rng.uniform(0, np.pi, 211), rng.uniform(0, 2*np.pi, 211)
    th[0], t[0] = 0.0, 0.0
    th[1] = np.pi
    R = D.RadialInterpolator(M, n)
    v = R(D.to_dev(fr), D.to_dev(th), D.to_dev(t)).cpu().numpy()
    ref = interpolate_radial_to_points(fr, th, t)
    assert np.abs(v - ref).max() < 2e-13*max(1.0, np.abs(ref).max())


def test_k0(D):
    from ipde_b200._lib import lib
    D.ensure_k0_table()
    x = np.concatenate([np.geomspace(1e-8, 2.0, 2000), np.linspace(2.0, 60.0, 4000), [100.0, 300.0, 700.0]])
    out = D.empty((x.size,))
    xd = D.to_dev(x)
    lib().ipde_eval_k0(D.ptr(xd), D.ptr(out), x.size, D.stream())
    ref = scipy.special.k0(x)
    ok = ref > 0
    assert np.abs(out.cpu().numpy()[ok]/ref[ok] - 1).max() < 1e-14


@pytest.mark.parametrize('kind,k', [(0, 0.0), (1, 3.0)])
@pytest.mark.parametrize('ns,nt', [(50, 7), (700, 3000), (5000, 129)])
def test_layer_apply_scalar(D, kind, k, ns, nt):
    from oracle.solvers import layer_apply
    rng = np.random.default_rng(ns + nt)
    sx, sy, q = rng.uniform(-1, 1, ns), rng.uniform(-1, 1, ns), rng.standard_normal(ns)
    tx, ty = rng.uniform(-1, 1, nt) + 2.5, rng.uniform(-1, 1, nt)
    out = D.layer_apply_scalar(kind, k, D.to_dev(sx), D.to_dev(sy), D.to_dev(q), D.to_dev(tx), D.to_dev(ty))
    ref = layer_apply('laplace' if kind == 0 else 'mh', k, sx, sy, q, tx, ty)
    assert rel(out.cpu().numpy(), ref) < 1e-13
    out2 = D.layer_apply_scalar(kind, k, D.to_dev(sx), D.to_dev(sy), D.to_dev(q), D.to_dev(tx), D.to_dev(ty), out=out, accumulate=True)
    assert rel(out2.cpu().numpy(), 2*ref) < 1e-13


def test_layer_apply_stokes(D):
    rng = np.random.default_rng(5)
    ns, nt = 300, 500
    sx, sy = rng.uniform(-1, 1, ns), rng.uniform(-1, 1, ns)
    fx, fy = rng.standard_normal(ns), rng.standard_normal(ns)
    tx, ty = rng.uniform(-1, 1, nt) + 2.5, rng.uniform(-1, 1, nt)
    out = D.layer_apply_stokes(*[D.to_dev(a) for a in (sx, sy, fx, fy, tx, ty)]).cpu().numpy()
    dx, dy = tx[:, None] - sx, ty[:, None] - sy
    r2 = dx*dx + dy*dy
    L = -0.5*np.log(r2)
    rf = (dx*fx + dy*fy)/r2
    u = ((L*fx + rf*dx).sum(1))/(4*np.pi)
    v = ((L*fy + rf*dy).sum(1))/(4*np.pi)
    p = rf.sum(1)/(2*np.pi)
    assert rel(out[0], u) < 1e-13 and rel(out[1], v) < 1e-13 and rel(out[2], p) < 1e-13


@pytest.mark.parametrize('M', [5, 64, 200, 1000])
def test_gemv_and_lu_solve(D, M):
    rng = np.random.default_rng(M)
    A = rng.standard_normal((M, M)) + M*np.eye(M)*0.1
    x = rng.standard_normal(M)
    y = D.gemv(D.to_dev(A), D.to_dev(x)).cpu().numpy()
    assert rel(y, A.dot(x)) < 1e-13
    B = rng.standard_normal((M, 2*M + 3))
    z = rng.standard_normal(2*M + 3)
    assert rel(D.gemv(D.to_dev(B), D.to_dev(z)).cpu().numpy(), B.dot(z)) < 1e-13
    lu = scipy.linalg.lu_factor(A)
    sol = D.DenseLU(lu).solve(D.to_dev(y)).cpu().numpy()
    ref = scipy.linalg.lu_solve(lu, y)
    assert rel(sol, ref) < 1e-11
    assert rel(A.dot(sol), y) < 1e-12


def test_lu_solve_ill_conditioned_is_backward_stable(D):
    # smooth-kernel matrix like the QFS check<-source map: tiny singular values, benign right-hand sides
    n = 400
    t = 2*np.pi*np.arange(n)/n
    src = 1.3*np.exp(1j*t)
    chk = 1.0*np.exp(1j*t)
    E = -np.log(np.abs(chk[:, None] - src[None, :]))/(2*np.pi)*(2*np.pi*1.3/n)
    rhs = np.exp(np.cos(t))*np.sin(2*t)
    lu = scipy.linalg.lu_factor(E)
    sol = D.DenseLU(lu).solve(D.to_dev(rhs)).cpu().numpy()
    assert np.abs(E.dot(sol) - rhs).max() < 1e-11*max(1.0, np.abs(sol).max())


def test_csr_spmv(D):
    from ipde_b200._lib import lib
    rng = np.random.default_rng(11)
    blocks = [rng.standard_normal((6, 6)) for _ in range(50)]
    A = scipy.sparse.block_diag(blocks, 'csr')
    x = rng.standard_normal(A.shape[1])
    y = D.empty((A.shape[0],))
    ip, ix = D.to_dev(A.indptr.astype(np.int32)), D.to_dev(A.indices.astype(np.int32))
    dd, xd = D.to_dev(A.data), D.to_dev(x)      # keep the device buffers alive across the call
    lib().ipde_csr_spmv_d(D.ptr(ip), D.ptr(ix), D.ptr(dd), D.ptr(xd), D.ptr(y), A.shape[0], D.stream())
    assert rel(y.cpu().numpy(), A.dot(x)) < 1e-14
    Az = A.astype(complex)*(1 + 0.5j)
    xz = x + 1j*rng.standard_normal(x.size)
    import torch
    yz = torch.empty(A.shape[0], dtype=torch.complex128, device='cuda')
    dz, xzd = D.to_dev(Az.data), D.to_dev(xz)
    lib().ipde_csr_spmv_z(D.ptr(ip), D.ptr(ix), D.ptr(dz), D.ptr(xzd), D.ptr(yz), A.shape[0], D.stream())
    assert rel(yz.cpu().numpy(), Az.dot(xz)) < 1e-14


def test_fd4_and_cubic_interp(D):
    from ipde_b200._lib import lib
    from oracle.solvers import fd_x_4, fd_y_4, lagrange_interp2d_periodic
    rng = np.random.default_rng(2)
    nx, ny, xh, yh = 40, 36, 0.1, 0.12
    f = rng.standard_normal((nx, ny))
    g = D.GridPlan(nx, ny, xh, yh)
    fx, fy = g.new_real(), g.new_real()
    fd = D.to_dev(f)
    lib().ipde_grid_fd4(g.h, D.ptr(fd), D.ptr(fx), D.ptr(fy), D.stream())
    assert rel(fx.cpu().numpy(), fd_x_4(f, xh)) < 1e-14
    assert rel(fy.cpu().numpy(), fd_y_4(f, yh)) < 1e-14
    px, py = rng.uniform(-1, 3, 100), rng.uniform(-1, 3, 100)
    out = D.empty((100,))
    pxd, pyd = D.to_dev(px), D.to_dev(py)
    lib().ipde_grid_interp_cubic(g.h, D.ptr(fd), -1.0, -1.0, D.ptr(pxd), D.ptr(pyd), 100, D.ptr(out), D.stream())
    ref = lagrange_interp2d_periodic(f, -1.0, -1.0, xh, yh, px, py, 3)
    assert rel(out.cpu().numpy(), ref) < 1e-13


def test_annular_against_reference_golden(D, golden_dir):
    from ipde_b200.annular import (ApproximateAnnularGeometry, RealAnnularGeometry,
                                   AnnularModifiedHelmholtzSolver, AnnularPoissonSolver)
    g = np.load(os.path.join(golden_dir, 'annular_scalar.npz'))
    n = int(g['n'])
    AAG = ApproximateAnnularGeometry(n, int(g['M']), float(g['width']), float(g['approx_r']))
    RAG = RealAnnularGeometry(g['speed'], g['curvature'], AAG)
    S = AnnularModifiedHelmholtzSolver(AAG, float(g['mh_k']))
    S.RAG = RAG
    v = g['mh_apply_in']
    assert rel(S._apply(v), g['mh_apply_out']) < 1e-13
    assert rel(S._optim_preconditioner(v), g['mh_prec_out']) < 1e-11
    u = S.solve(RAG, g['mh_f'], g['mh_ig'], g['mh_og'], tol=1e-14, maxiter=10, restart=50)
    assert np.abs(u - g['mh_u']).max() < 1e-12
    assert S.iterations_last_call < 40
    u0 = S.solve(RAG, g['mh_f'], None, None, tol=1e-14, maxiter=10, restart=50)
    assert np.abs(u0 - g['mh_u_zero_bc']).max() < 1e-12
    P = AnnularPoissonSolver(AAG)
    up = P.solve(RAG, g['po_f'], g['mh_ig'], g['mh_og'], tol=1e-14, maxiter=10, restart=50)
    assert np.abs(up - g['po_u']).max() < 1e-12


def test_annular_restart_and_iteration_parity_with_oracle(D):
    from ipde_b200 import annular as A
    from oracle import annular as O
    n, M, width = 200, 12, 0.12
    t = np.linspace(0, 2*np.pi, n, endpoint=False)
    c = (1 + 0.2*np.cos(3*t))*np.exp(1j*t)
    from ipde_b200.geometry import GSB
    b = GSB(c=c)
    rng = np.random.default_rng(0)
    f = rng.standard_normal((M, n))
    kw = dict(tol=1e-12, restart=8, maxiter=30)
    Ao = O.ApproximateAnnularGeometry(n, M, width, 1.0)
    So = O.AnnularModifiedHelmholtzSolver(Ao, 4.0)
    uo = So.solve(O.RealAnnularGeometry(b.speed, b.curvature, Ao), f, np.zeros(n), np.zeros(n), **kw)
    Ad = A.ApproximateAnnularGeometry(n, M, width, 1.0)
    Sd = A.AnnularModifiedHelmholtzSolver(Ad, 4.0)
    ud = Sd.solve(A.RealAnnularGeometry(b.speed, b.curvature, Ad), f, None, None, **kw)
    assert rel(ud, uo) < 1e-10
    assert abs(Sd.iterations_last_call - So.iterations_last_call) <= 2


# ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('kind,kappa,N', [(0, 0.0, 96), (1, 2.0, 96), (1, 5.0, 160), (1, 0.3, 128)])
def test_free_space_evaluator_matches_direct_sum(D, kind, kappa, N):
    """Spectral-Ewald evaluator (csrc/ewald.cu) vs the all-pairs sum (csrc/layer.cu) and vs NumPy."""
    import scipy.special as ss
    import torch
    h = 2.4/N
    x0 = -1.2
    rng = np.random.default_rng(3)
    ns = 300
    th = np.sort(rng.uniform(0, 2*np.pi, ns))
    sx, sy = 0.8*np.cos(th) + 0.013, 0.75*np.sin(th) - 0.007
    q = rng.standard_normal(ns)
    xv = x0 + h*np.arange(N)
    X, Y = np.meshgrid(xv, xv, indexing='ij')
    sel = (X/0.8)**2 + (Y/0.75)**2 < 0.93**2
    gix, giy = np.nonzero(sel)
    px_, py_ = 0.74*np.cos(th[::2]) + 0.001, 0.7*np.sin(th[::2])
    ev = D.FreeSpaceEvaluator(kind, kappa, x0, x0, h, sx, sy, grid_nodes=(gix, giy),
                              grid_out_pos=np.arange(gix.size), points=(px_, py_))
    out_grid = D.zeros((gix.size,))
    out_pts = ev(D.to_dev(q), grid_out=out_grid)
    G = (lambda r: -np.log(r)/(2*np.pi)) if kind == 0 else (lambda r: ss.k0(kappa*r)/(2*np.pi))
    tx = np.concatenate([X[sel], px_]); ty = np.concatenate([Y[sel], py_])
    exact = G(np.hypot(tx[:, None] - sx, ty[:, None] - sy)).dot(q)
    got = np.concatenate([out_grid.cpu().numpy(), out_pts.cpu().numpy()])
    scale = np.abs(exact).max()
    assert np.abs(got - exact).max()/scale < 2e-13
    # second application accumulates on the grid output and is deterministic
    out_grid2 = D.zeros((gix.size,))
    out_pts2 = ev(D.to_dev(q), grid_out=out_grid2)
    assert torch.equal(out_grid, out_grid2) and torch.equal(out_pts, out_pts2)


@pytest.mark.parametrize('kind', ['laplace', 'modified_helmholtz'])
@pytest.mark.parametrize('interior', [True, False])
def test_qfs_device_build_matches_host_build(D, kind, interior):
    """QFS operators built on the device (kernel matrices, FFT upsampling, cuSOLVER SVD) generate the same
    source strengths as the NumPy build for smooth densities."""
    from ipde_b200.geometry import GSB, star
    from ipde_b200.qfs import Laplace_QFS, Modified_Helmholtz_QFS
    bdy = GSB(c=star(240, a=0.15, f=5))
    mk = (lambda **kw: Laplace_QFS(bdy, interior, True, True, **kw)) if kind == 'laplace' else \
         (lambda **kw: Modified_Helmholtz_QFS(bdy, interior, True, True, 3.0, **kw))
    qh = mk(device_build=False)
    qd = mk(device_build=True)
    assert qd.on_device and not qh.on_device
    t = bdy.t
    tau = [np.exp(np.cos(t))*np.sin(2*t), np.cos(3*t + 0.3)]
    sh, sd = qh(tau), qd(tau)
    # strengths agree up to the truncated directions; what has to agree is the potential they generate
    th = np.linspace(0, 2*np.pi, 50, endpoint=False)
    rad = 0.5 if interior else 2.0
    tx, ty = rad*np.cos(th), rad*np.sin(th)
    ker = qh.kernel
    K = ker.slp(tx, ty, qh.source.x, qh.source.y)*qh.source.weights
    assert rel(K.dot(sd), K.dot(sh)) < 1e-12
    ub = np.exp(np.sin(t))
    assert rel(K.dot(qd.u2s(ub)), K.dot(qh.u2s(ub))) < 1e-11


@pytest.mark.parametrize('interior', [True, False])
def test_stokes_qfs_device_build_matches_host_build(D, interior):
    """Stokes QFS operators formed and factored on the device generate the same velocity field as the NumPy build"""
    from ipde_b200.geometry import GSB, star
    from ipde_b200.qfs import Stokes_QFS
    bdy = GSB(c=star(200, a=0.15, f=5))
    qh = Stokes_QFS(bdy, interior, True, True, device_build=False)
    qd = Stokes_QFS(bdy, interior, True, True, device_build=True)
    assert qd.on_device and not qh.on_device
    t = bdy.t
    tau = [np.concatenate([np.exp(np.cos(t))*np.sin(2*t), np.cos(t)]), np.concatenate([np.cos(3*t + 0.3), np.sin(2*t)])]
    sh, sd = qh(tau), qd(tau)
    th = np.linspace(0, 2*np.pi, 50, endpoint=False)
    rad = 0.5 if interior else 2.0
    tx, ty = rad*np.cos(th), rad*np.sin(th)
    K = qh.kernel.slp(tx, ty, qh.source.x, qh.source.y)*np.tile(qh.source.weights, 2)
    assert rel(K.dot(sd), K.dot(sh)) < 1e-11
    ub = np.concatenate([np.exp(np.sin(t)), np.cos(2*t)])
    assert rel(K.dot(qd.u2s(ub)), K.dot(qh.u2s(ub))) < 1e-10


def test_nufft2_apply3_matches_separate_transforms(D):
    """value + gradient from one fine-grid transform (kernel differentiation) vs three spectral transforms"""
    n = 96
    rng = np.random.default_rng(5)
    g = D.GridPlan(n, n, 2.4/n, 2.4/n)
    f = rng.standard_normal((n, n))
    spec = g.fft2(D.to_dev(f))
    x, y = rng.uniform(0, 2*np.pi, 300), rng.uniform(0, 2*np.pi, 300)
    nu = D.Nufft2(g, x, y)
    v, dx, dy = nu.apply3(spec)
    v0, dx0, dy0 = nu(spec, 0, 0), nu(spec, 1, 0), nu(spec, 0, 1)
    assert rel(v.cpu().numpy(), v0.cpu().numpy()) < 1e-13
    assert rel(dx.cpu().numpy(), dx0.cpu().numpy()) < 5e-13
    assert rel(dy.cpu().numpy(), dy0.cpu().numpy()) < 5e-13


def test_annular_stokes_against_reference_golden(D, golden_dir):
    """device annular Stokes operator, preconditioner and GMRES solve against the reference's own outputs"""
    from ipde_b200.annular import ApproximateAnnularGeometry, RealAnnularGeometry, AnnularStokesSolver
    g = np.load(os.path.join(golden_dir, 'annular_stokes.npz'))
    AAG = ApproximateAnnularGeometry(int(g['n']), int(g['M']), float(g['width']), float(g['radius']), drop_nyquist=True)
    RAG = RealAnnularGeometry(g['speed'], g['curvature'], AAG)
    S = AnnularStokesSolver(AAG, 1.0)
    S.RAG = RAG
    assert rel(S._apply_optim_real(g['x']), g['Ax']) < 1e-12
    assert rel(S._preconditioner(g['x']), g['Px']) < 1e-11
    ur, ut, p = S.solve(RAG, g['fr'], g['ft'], g['irg'], g['itg'], g['org'], g['otg'], tol=1e-14, restart=80, maxiter=4)
    assert rel(ur, g['ur']) < 1e-11 and rel(ut, g['ut']) < 1e-11 and rel(p, g['p']) < 1e-10


def test_stokes_free_space_evaluator_matches_direct_sum(D):
    """stokeslet velocity + pressure through the Laplace spectral-Ewald evaluator (value + gradient fields with
    affine weights) against the all-pairs stokeslet kernel (csrc/layer.cu) on grid nodes and off-grid points"""
    import torch
    N = 128
    h = 2.4/N
    x0 = -1.2
    rng = np.random.default_rng(7)
    ns = 300
    th = np.sort(rng.uniform(0, 2*np.pi, ns))
    sx, sy = 0.8*np.cos(th) + 0.013, 0.75*np.sin(th) - 0.007
    fx, fy = rng.standard_normal(ns), rng.standard_normal(ns)
    xv = x0 + h*np.arange(N)
    X, Y = np.meshgrid(xv, xv, indexing='ij')
    sel = (X/0.8)**2 + (Y/0.75)**2 < 0.92**2
    gix, giy = np.nonzero(sel)
    px_, py_ = 0.72*np.cos(th[::2]) + 0.001, 0.68*np.sin(th[::2])
    ev = D.StokesFreeSpaceEvaluator(x0, x0, h, sx, sy, grid_nodes=(gix, giy), grid_out_pos=np.arange(gix.size),
                                    points=(px_, py_))
    ng = gix.size
    outs = (D.zeros((ng,)), D.zeros((ng,)), D.zeros((ng,)))
    dfx, dfy = D.to_dev(fx), D.to_dev(fy)
    pts = ev(dfx, dfy, grid_out=outs)
    tx = np.concatenate([X[sel], px_]); ty = np.concatenate([Y[sel], py_])
    ref = D.layer_apply_stokes(D.to_dev(sx), D.to_dev(sy), dfx, dfy, D.to_dev(tx), D.to_dev(ty)).cpu().numpy()
    got = np.concatenate([np.stack([o.cpu().numpy() for o in outs]), pts.cpu().numpy()], axis=1)
    for k, name in enumerate(('u', 'v', 'p')):
        err = np.abs(got[k] - ref[k]).max()/np.abs(ref[k]).max()
        assert err < 5e-11, (name, err)


@pytest.mark.parametrize('kind', ['laplace', 'modified_helmholtz'])
@pytest.mark.parametrize('n', [1200, 4096])
def test_structured_qfs_device_apply(D, kind, n):
    """k_banded_apply + FFTs (DeviceStructuredQFS) against the NumPy statement of the same operator, and the
    potential of the device-computed sources against Green's formula half a node spacing from the curve."""
    import torch
    from ipde_b200.geometry import GSB, star
    from ipde_b200 import qfs as Q
    from ipde_b200.solvers.internals.scalar import device_qfs
    from test_host import _green_case
    k = 3.0 if kind == 'modified_helmholtz' else 0.0
    bdy = GSB(c=star(n, a=0.15, f=5))
    mk = (lambda i: Q.Laplace_QFS(bdy, i, True, True, tol=1e-13, structured=True)) if k == 0 else \
         (lambda i: Q.Modified_Helmholtz_QFS(bdy, i, True, True, k, tol=1e-13, structured=True))
    for interior in (True, False):
        q = mk(interior)
        dq = device_qfs(q)
        rng = np.random.default_rng(5)
        t = bdy.t
        tau = np.concatenate([np.exp(np.cos(t))*np.sin(2*t) + 1e-6*rng.standard_normal(n), np.cos(3*t + 0.3)])
        sd = dq(D.to_dev(tau)).cpu().numpy()
        sh = q([tau[:n], tau[n:]])
        # the strengths amplify the top density modes by exp(alpha_s pi) ~ 1e6: the round-off (1e-16) of the two
        # different FFTs in those modes shows up at ~5e-9 of max|sigma|, not in the field sigma generates
        th = np.linspace(0, 2*np.pi, 97, endpoint=False)
        z = (1 + 0.15*np.cos(5*th))*np.exp(1j*th)*(1 - 2*np.pi/n if interior else 1 + 2*np.pi/n)
        field = lambda sg: Q.layer_apply_matrix_free(q.kernel, q.source, z.real, z.imag, sg)
        assert rel(sd, sh) < 1e-7
        dfield = np.abs(field(sd) - field(sh)).max()/max(1.0, np.abs(field(sh)).max())
        assert dfield < 1e-12, dfield
        ub = np.exp(np.sin(t))
        su, suh = dq.u2s(D.to_dev(ub)).cpu().numpy(), q.u2s(ub)
        assert rel(su, suh) < 1e-7
        dfield = np.abs(field(su) - field(suh)).max()/max(1.0, np.abs(field(suh)).max())
        assert dfield < 1e-12, dfield

        class OnDevice(object):          # the same Green-identity check, sources computed by the device apply
            kernel, source = q.kernel, q.source
            def __call__(self, taus):
                return dq(D.to_dev(np.concatenate(taus))).cpu().numpy()
            def u2s(self, u):
                return dq.u2s(D.to_dev(u)).cpu().numpy()
        val, val2, ex = _green_case(bdy, interior, k, OnDevice(), 0.5)
        assert np.abs(val - ex).max() < 1e-12*max(1.0, np.abs(ex).max())
        assert np.abs(val2 - ex).max() < 1e-12*max(1.0, np.abs(ex).max())


@pytest.mark.parametrize('shape', [(96, 80), (1024, 48), (2048, 64), (474, 94)])
def test_grid_solve_stokes_and_stress(D, shape):
    """k_stokes_multiply / k_stokes_stress against the reference's own arithmetic in NumPy:
    StokesSolver._grid_solve (ipde/solvers/multi_boundary/stokes.py:34-50) and the stress assembly of
    VectorSolver.__call__ with Fourier derivatives (multi_boundary/vector.py:68-77, derivatives.py:25-28)."""
    from ipde_b200._lib import lib
    nx, ny = shape
    xh, yh = 2.4/nx, 2.0/ny
    rng = np.random.default_rng(11)
    fu, fv = rng.standard_normal(shape), rng.standard_normal(shape)
    bumpy = np.abs(rng.standard_normal(shape))
    bumpy /= bumpy.sum()*xh*yh
    g = D.GridPlan(nx, ny, xh, yh)
    d = D.to_dev
    fuc, fvc, bd = d(fu), d(fv), d(bumpy)
    uh, vh, ph, work = g.new_spec(), g.new_spec(), g.new_spec(), g.new_spec()
    uc, vc, pc = g.new_real(), g.new_real(), g.new_real()
    P = D.ptr
    lib().ipde_grid_solve_stokes(g.h, P(fuc), P(fvc), P(bd), P(uh), P(vh), P(ph), P(work), P(uc), P(vc), P(pc), D.stream())
    kx = np.fft.fftfreq(nx, xh/(2*np.pi))[:, None]
    ky = np.fft.fftfreq(ny, yh/(2*np.pi))
    ikx, iky = 1j*kx, 1j*ky
    lap = -kx*kx - ky*ky
    lap[0, 0] = np.inf
    ilap = 1.0/lap
    dm = lambda f: f - f.sum()*xh*yh*bumpy
    fuh, fvh = np.fft.fft2(dm(fu)), np.fft.fft2(dm(fv))
    cph = ilap*(ikx*fuh + iky*fvh)
    cuh = ilap*(ikx*cph - fuh)
    cvh = ilap*(iky*cph - fvh)
    ru, rv, rp = (np.fft.ifft2(x).real for x in (cuh, cvh, cph))
    for got, ref in ((uc, ru), (vc, rv), (pc, rp)):
        assert rel(got.cpu().numpy(), ref) < 1e-11
    # stress spectra -> back to the grid; reference: fourier derivatives of the real fields
    txx, txy, tyy = g.new_spec(), g.new_spec(), g.new_spec()
    lib().ipde_grid_stokes_stress(g.h, P(uh), P(vh), P(ph), P(txx), P(txy), P(tyy), D.stream())
    dx = lambda f: np.fft.ifft2(np.fft.fft2(f)*ikx).real
    dy = lambda f: np.fft.ifft2(np.fft.fft2(f)*iky).real
    refs = (2*dx(ru) - rp, dy(ru) + dx(rv), 2*dy(rv) - rp)
    for spec, ref in zip((txx, txy, tyy), refs):
        assert rel(g.ifft2(spec).cpu().numpy(), ref) < 1e-11
