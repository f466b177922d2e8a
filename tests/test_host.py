"""CPU: host-side setup code (geometry, cutoffs, QFS) and the C-ABI surface."""
import os
import re
import numpy as np
import pytest
from ipde_b200.geometry import GSB, star, fourier_resample
from ipde_b200.slepian import SlepianMollifier
from ipde_b200 import qfs as Q
from problems import single_star, three_bodies


def test_slepian_matches_reference_tables(golden_dir):
    g = np.load(os.path.join(golden_dir, 'slepian_golden.npz'))
    x = g['x']
    for r in [4, 8, 16, 24, 30, 32, 40, 60]:
        S = SlepianMollifier(r)
        assert np.abs(S.step(x) - g['step_%d' % r]).max() < 2e-14
        assert np.abs(S.bump(x) - g['bump_%d' % r]).max() < 2e-14


def test_gsb_circle():
    b = GSB(c=1.5*np.exp(1j*np.linspace(0, 2*np.pi, 64, endpoint=False)))
    assert np.abs(b.curvature - 1/1.5).max() < 1e-13
    assert np.abs(b.normal_x - np.cos(b.t)).max() < 1e-13
    assert abs(b.perimeter - 3*np.pi) < 1e-12


def test_fourier_resample_roundtrip():
    t = np.linspace(0, 2*np.pi, 32, endpoint=False)
    f = np.exp(np.sin(t)) + 0.1*np.cos(16*t)
    up = fourier_resample(f, 96)
    assert np.abs(up[::3] - f).max() < 1e-13
    assert np.abs(fourier_resample(up, 32) - f).max() < 1e-13


def test_local_coordinates_identity():
    ebdyc = single_star(nb=200, M=8)
    e = ebdyc[0]
    bdy = e.bdy
    k = np.fft.fftfreq(bdy.N, 1.0/bdy.N)
    ch = np.fft.fft(bdy.c)
    E = np.exp(1j*np.outer(e.grid_ia_t, k))
    X = E.dot(ch)/bdy.N
    Xp = E.dot(ch*1j*k)/bdy.N
    nrm = -1j*Xp/np.abs(Xp)
    z = X + e.grid_ia_r*nrm
    assert np.abs(z - (e.grid_ia_x + 1j*e.grid_ia_y)).max() < 1e-12
    assert e.grid_ia_r.max() <= 0 and e.grid_ia_r.min() >= -e.radial_width


def test_collection_bookkeeping():
    ebdyc = three_bodies(nb=60, M=6)
    assert ebdyc.phys.sum() == ebdyc.grid_phys.N
    assert ebdyc.grid_pna.N + sum(len(e.grid_ia_xind) for e in ebdyc) == ebdyc.grid_phys.N
    assert ebdyc.grid_step.max() <= 1.0 + 1e-14
    # every annulus point is physical
    for e in ebdyc:
        assert ebdyc.phys[e.grid_ia_xind, e.grid_ia_yind].all()


@pytest.mark.parametrize('interior', [True, False])
def test_qfs_green_identity_modified_helmholtz(interior):
    import scipy.special as ss
    k = 3.0
    bdy = GSB(c=star(300, a=0.15, f=5))
    tt = np.linspace(0, 2*np.pi, 333, endpoint=False)
    q = Q.Modified_Helmholtz_QFS(bdy, interior, True, True, k, tol=1e-13)
    if interior:
        u = lambda x, y: np.exp(k*(0.6*x + 0.8*y))/10
        un = k*u(bdy.x, bdy.y)*(0.6*bdy.normal_x + 0.8*bdy.normal_y)
        sig = q([un, -u(bdy.x, bdy.y)])
        z = (1 + 0.15*np.cos(5*tt))*np.exp(1j*tt)*0.995
        ex = u(z.real, z.imag)
    else:
        x0, y0 = 0.1, 0.2
        r = np.hypot(bdy.x - x0, bdy.y - y0)
        ue = ss.k0(k*r)/(2*np.pi)
        uen = -k*ss.k1(k*r)/(2*np.pi)*((bdy.x - x0)*bdy.normal_x + (bdy.y - y0)*bdy.normal_y)/r
        sig = q([-uen, ue])
        z = (1 + 0.15*np.cos(5*tt))*np.exp(1j*tt)*1.005
        ex = ss.k0(k*np.hypot(z.real - x0, z.imag - y0))/(2*np.pi)
    val = Q.layer_apply_matrix_free(q.kernel, q.source, z.real, z.imag, sig)
    assert np.abs(val - ex).max() < 5e-12


def test_c_abi_exports_every_declared_symbol():
    from ipde_b200._lib import lib, parse_header, LIB_PATH
    protos = parse_header()
    assert len(protos) >= 45
    L = lib()         # raises if a declared symbol is missing from the .so
    assert L.ipde_version() >= 100
    import subprocess
    out = subprocess.run(['nm', '-D', '--defined-only', LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r' T (ipde_\w+)', out))
    assert set(protos) <= exported
    # nothing torch-typed in the ABI
    assert 'torch' not in open(os.path.join(os.path.dirname(LIB_PATH), '..', 'include', 'ipde_b200.h')).read()


def test_annulus_points_outside_the_physical_mask_are_flagged():
    """A grid point within round-off of the boundary can get annular coordinates and still fail the inside test; it
    has no slot in the compact vector and must be excluded from compact scatters (ia_is_phys).  generate_grid pads
    automatically chosen sizes to lengths the device FFT supports."""
    from problems import single_star, three_bodies
    from ipde_b200._lib import lib
    e = single_star(nb=400, M=12, ng=480)
    eb = e[0]
    assert eb.ia_is_phys.shape == eb.ia_inds.shape
    n_out = int((~eb.ia_is_phys).sum())
    assert n_out >= 1
    assert int(e.phys.sum()) == int(e.phys_not_in_annulus.sum()) + int(eb.ia_is_phys.sum())
    # every physical in-annulus point has its own slot; the flagged ones map to the placeholder slot 0
    slots = eb.ia_inds[eb.ia_is_phys]
    assert np.unique(slots).size == slots.size and np.all(eb.ia_inds[~eb.ia_is_phys] == 0)
    g = three_bodies(nb=200, M=12).grid
    assert lib().ipde_fft_size_supported(g.Nx) and lib().ipde_fft_size_supported(g.Ny//2)


def _green_case(bdy, interior, k, q, dist):
    """densities of a known PDE solution on the curve -> sources; returns (potential of the sources, exact) on a
    test curve `dist` node spacings into the evaluation side, and the same for u2s"""
    import scipy.special as ss
    tt = np.linspace(0, 2*np.pi, 333, endpoint=False)
    N = bdy.N
    E = np.exp(1j*np.outer(tt, np.fft.fftfreq(N, 1.0/N)))
    ch = np.fft.fft(bdy.c)/N
    Z = E.dot(ch)
    Zp = E.dot(ch*1j*np.fft.fftfreq(N, 1.0/N))
    z = Z + (-1j*Zp/np.abs(Zp))*(-dist*bdy.dt*np.abs(Zp) if interior else dist*bdy.dt*np.abs(Zp))
    if interior:
        u = (lambda x, y: np.exp(k*(0.6*x + 0.8*y))/10) if k > 0 else (lambda x, y: np.exp(0.6*x)*np.cos(0.6*y))
        eps = 1e-6
        ub = u(bdy.x, bdy.y)
        if k > 0:
            un = k*ub*(0.6*bdy.normal_x + 0.8*bdy.normal_y)
        else:
            un = 0.6*np.exp(0.6*bdy.x)*(np.cos(0.6*bdy.y)*bdy.normal_x - np.sin(0.6*bdy.y)*bdy.normal_y)
        sig = q([un, -ub])
        ex = u(z.real, z.imag)
    else:
        x0, y0 = 0.1, 0.2
        dx, dy = bdy.x - x0, bdy.y - y0
        r = np.hypot(dx, dy)
        if k > 0:
            ub = ss.k0(k*r)/(2*np.pi)
            un = -k*ss.k1(k*r)/(2*np.pi)*(dx*bdy.normal_x + dy*bdy.normal_y)/r
            ex = ss.k0(k*np.hypot(z.real - x0, z.imag - y0))/(2*np.pi)
        else:
            ub = dx/r**2
            un = (1/r**2 - 2*dx*dx/r**4)*bdy.normal_x + (-2*dx*dy/r**4)*bdy.normal_y
            ex = (z.real - x0)/((z.real - x0)**2 + (z.imag - y0)**2)
        sig = q([-un, ub])
    val = Q.layer_apply_matrix_free(q.kernel, q.source, z.real, z.imag, sig)
    val2 = Q.layer_apply_matrix_free(q.kernel, q.source, z.real, z.imag, q.u2s(ub))
    return val, val2, ex


@pytest.mark.parametrize('interior', [True, False])
@pytest.mark.parametrize('k', [0.0, 3.0])
def test_structured_qfs_green_identity(interior, k):
    """Structured (Fourier-banded) QFS: Green's representation formula half a node spacing away from the curve and
    the u2s map, at 1e-12 (the reference's qfs objects are absent; this identity is what pins the restatement)."""
    from ipde_b200.qfs_structured import StructuredQFS
    bdy = GSB(c=star(1200, a=0.15, f=5))
    ker = Q.ModifiedHelmholtzKernel(k) if k > 0 else Q.LaplaceKernel()
    q = StructuredQFS(bdy, interior, True, True, ker, tol=1e-13, device='cpu')
    assert q.op.BH < 400 and q.source.N == q.M
    val, val2, ex = _green_case(bdy, interior, k, q, 0.5)
    assert np.abs(val - ex).max() < 1e-12*max(1.0, np.abs(ex).max())
    assert np.abs(val2 - ex).max() < 1e-12*max(1.0, np.abs(ex).max())


def test_structured_qfs_matches_dense_qfs():
    """same densities through the dense SVD-factored operators and through the structured ones: the potentials on
    the evaluation side agree to 1e-12 (the source curves differ, the represented field does not)"""
    from ipde_b200.qfs_structured import StructuredQFS
    bdy = GSB(c=star(600, a=0.1, f=3))
    k = 2.0
    ker = Q.ModifiedHelmholtzKernel(k)
    qs = StructuredQFS(bdy, True, True, True, ker, tol=1e-13, device='cpu')
    qd = Q.QFS(bdy, True, True, True, ker, tol=1e-13, device_build=False)
    t = bdy.t
    tau = [np.exp(np.cos(t))*np.sin(2*t), np.cos(3*t + 0.3)]
    th = np.linspace(0, 2*np.pi, 200, endpoint=False)
    for fac in (0.3, 0.9, 0.995):
        z = (1 + 0.1*np.cos(3*th))*np.exp(1j*th)*fac
        vs = Q.layer_apply_matrix_free(ker, qs.source, z.real, z.imag, qs(tau))
        vd = Q.layer_apply_matrix_free(ker, qd.source, z.real, z.imag, qd(tau))
        assert np.abs(vs - vd).max() < 1e-12


def test_registration_torch_path_matches_numpy_path(monkeypatch):
    """near.py: the tensor (GPU) statement of the near-curve search, Newton coordinates and inside test, run here on
    CPU tensors, against the NumPy statement -- identical point sets in identical order, coordinates to round-off"""
    import torch
    import ipde_b200.near as near
    from problems import four_bodies
    a = four_bodies(nb=120, M=8)
    monkeypatch.setattr(near, '_torch_device', lambda n: torch.device('cpu'))
    b = four_bodies(nb=120, M=8)
    assert np.array_equal(a.phys, b.phys)
    assert np.abs(a.grid_step - b.grid_step).max() < 1e-12
    for ea, eb in zip(a, b):
        assert np.array_equal(ea.grid_ia_xind, eb.grid_ia_xind) and np.array_equal(ea.grid_ia_yind, eb.grid_ia_yind)
        assert np.abs(ea.grid_ia_r - eb.grid_ia_r).max() < 1e-13
        assert np.abs(np.angle(np.exp(1j*(ea.grid_ia_t - eb.grid_ia_t)))).max() < 1e-13


def test_registration_against_brute_force_closest_point():
    """near.py (the restatement of near_finder): for a sample of annulus grid points |r| equals the brute-force
    distance to a 64x upsampled polygon of the curve, the sign of r matches a winding-number inside test, and t is the
    parameter of the closest point"""
    ebdyc = three_bodies(nb=80, M=8)
    rng = np.random.default_rng(3)
    for e in ebdyc:
        bdy = e.bdy
        fine = fourier_resample(bdy.c, 64*bdy.N)
        tf = 2*np.pi*np.arange(fine.size)/fine.size
        sel = rng.choice(e.grid_ia_xind.size, size=min(150, e.grid_ia_xind.size), replace=False)
        z = ebdyc.grid.xv[e.grid_ia_xind[sel]] + 1j*ebdyc.grid.yv[e.grid_ia_yind[sel]]
        d = np.abs(z[:, None] - fine[None, :])
        j = d.argmin(axis=1)
        hmax = np.abs(np.diff(fine)).max()
        ra = np.abs(e.grid_ia_r[sel])
        # the nearest polygon vertex is at most s = hmax/2 along the curve from the foot point:
        # d^2 <= r^2 + s^2 (1 + kappa r) + O(s^3)
        bend = 1 + np.abs(bdy.curvature).max()*ra
        bound = np.minimum(0.5*hmax, bend*hmax**2/(8*np.maximum(ra, 1e-300))) + 1e-12
        assert np.all(d.min(axis=1) - ra >= -1e-12) and np.all(d.min(axis=1) - ra <= 1.05*bound)
        assert np.abs(np.angle(np.exp(1j*(e.grid_ia_t[sel] - tf[j])))).max() < 1.5*(tf[1] - tf[0])
        # winding number: inside <-> r < 0 for a counter-clockwise curve
        ang = np.angle((np.roll(fine, -1)[None, :] - z[:, None])/(fine[None, :] - z[:, None])).sum(axis=1)/(2*np.pi)
        inside = np.abs(ang) > 0.5
        assert np.array_equal(inside, e.grid_ia_r[sel] < 0)


def test_point_partition_against_brute_force():
    """EmbeddedPointPartition (ipde/ebdy_collection.py:37-229): zones of random points against a brute-force
    classification (distance to a 64x upsampled polygon + winding number), coordinates against the closest vertex"""
    from ipde_b200.ebdy_collection import EmbeddedPointPartition
    ebdyc = three_bodies(nb=80, M=8)
    g = ebdyc.grid
    rng = np.random.default_rng(5)
    n = 3000
    x = rng.uniform(g.x_bounds[0], g.x_bounds[1], n)
    y = rng.uniform(g.y_bounds[0], g.y_bounds[1], n)
    p = EmbeddedPointPartition(ebdyc, x, y)
    assert p.zone1_N + p.zone2_N + p.zone3_N >= n          # a point may be near two boundaries
    z = x + 1j*y
    in_any = np.zeros(n, dtype=bool)
    flagged = np.zeros(n, dtype=bool)
    for ind, e in enumerate(ebdyc):
        fine = fourier_resample(e.bdy.c, 64*e.bdy.N)
        d = np.abs(z[:, None] - fine[None, :]).min(axis=1)
        ang = np.angle((np.roll(fine, -1)[None, :] - z[:, None])/(fine[None, :] - z[:, None])).sum(axis=1)/(2*np.pi)
        inside_curve = np.abs(ang) > 0.5
        physical_side = inside_curve if e.interior else ~inside_curve
        w = e.radial_width
        margin = 1e-3*w                                     # polygon resolution
        sure_annulus = physical_side & (d < w - margin)
        sure_exterior = ~physical_side & (d < w - margin) & (d > margin)
        sure_far = d > w + margin
        z2 = np.zeros(n, dtype=bool); z2[p.zone2l[ind]] = True
        z3 = np.zeros(n, dtype=bool); z3[p.zone3l[ind]] = True
        assert np.all(z2[sure_annulus]) and not np.any(z2[sure_exterior | sure_far])
        # zone 3: every located point on the non-physical side (the search radius is a little more than the width)
        assert np.all(z3[sure_exterior]) and not np.any(z3 & physical_side & (d > margin)) and not np.any(z3[d > 2*w])
        # coordinates: |r| is the distance to the curve, t the parameter of the closest point
        near = p.zone2l[ind]
        ra = np.abs(p.zone2r[ind])
        hmax = np.abs(np.diff(fine)).max()
        bound = np.minimum(0.5*hmax, (1 + np.abs(e.bdy.curvature).max()*ra)*hmax**2/(8*np.maximum(ra, 1e-300))) + 1e-12
        assert np.all(d[near] - ra >= -1e-12) and np.all(d[near] - ra <= 1.05*bound)
        tf = 2*np.pi*np.arange(fine.size)/fine.size
        j = np.abs(z[near, None] - fine[None, :]).argmin(axis=1)
        assert np.abs(np.angle(np.exp(1j*(p.zone2t[ind] - tf[j])))).max() < 3*(tf[1] - tf[0])
        assert np.all((p.zone2_transfr[ind] >= 0) & (p.zone2_transfr[ind] <= np.pi))
        in_any |= z2
        flagged |= z2 | z3
    assert np.array_equal(p.zone1, ~flagged) or np.all(p.zone1[~flagged])
    assert not np.any(p.zone1 & in_any)
    assert p.x_transf.size == p.zone1_N and p.x_transf.min() >= 0 and p.x_transf.max() <= 2*np.pi


@pytest.mark.parametrize('kind', ['laplace', 'modified_helmholtz'])
@pytest.mark.parametrize('interior', [True, False])
def test_qfs_against_layer_potential_quadrature(kind, interior):
    """QFS (the restatement of the absent `qfs` package) independently of Green's formula: for arbitrary smooth
    densities the effective sources reproduce the single + double layer potential of the boundary at test points well
    away from it, where the periodic trapezoid rule on a 4x upsampled curve is spectrally accurate"""
    import scipy.special as ss
    k = 2.5
    bdy = GSB(c=star(240, a=0.12, f=4))
    t = bdy.t
    sigma = np.exp(np.cos(t))*np.sin(2*t) + 0.3
    mu = np.cos(3*t + 0.4) - 0.5*np.sin(t)
    if kind == 'laplace':
        q = Q.Laplace_QFS(bdy, interior, True, True, tol=1e-13)
    else:
        q = Q.Modified_Helmholtz_QFS(bdy, interior, True, True, k, tol=1e-13)
    src = q([sigma, mu])
    tt = np.linspace(0, 2*np.pi, 41, endpoint=False)
    z = (0.45 if interior else 1.9)*np.exp(1j*tt)*(1 + 0.1*np.cos(2*tt))
    val = Q.layer_apply_matrix_free(q.kernel, q.source, z.real, z.imag, src)
    # quadrature of the layer potentials on the upsampled curve
    up = 4
    fine = GSB(c=fourier_resample(bdy.c, up*bdy.N))
    sf, mf = fourier_resample(sigma, up*bdy.N).real, fourier_resample(mu, up*bdy.N).real
    dx, dy = z.real[:, None] - fine.x[None, :], z.imag[:, None] - fine.y[None, :]
    r = np.hypot(dx, dy)
    dn = dx*fine.normal_x[None, :] + dy*fine.normal_y[None, :]          # (x - y) . n_y
    if kind == 'laplace':
        S = -np.log(r)/(2*np.pi)
        Dk = dn/(2*np.pi*r*r)
    else:
        S = ss.k0(k*r)/(2*np.pi)
        Dk = k*ss.k1(k*r)*dn/r/(2*np.pi)
    ref = (S*(sf*fine.weights)[None, :]).sum(axis=1) + (Dk*(mf*fine.weights)[None, :]).sum(axis=1)
    assert np.abs(val - ref).max() < 2e-12*max(1.0, np.abs(ref).max()), np.abs(val - ref).max()


@pytest.mark.parametrize('interior', [True, False])
def test_stokes_qfs_against_layer_potential_quadrature(interior):
    """the same for the Stokes QFS: stokeslet + stresslet layer potentials of arbitrary smooth vector densities (with
    a net normal component: its single layer carries no velocity and is the direction the density -> source map drops)
    against the trapezoid rule on the upsampled curve, velocity at far test points"""
    bdy = GSB(c=star(240, a=0.12, f=4))
    t = bdy.t
    sig = np.concatenate([np.exp(np.cos(t))*np.sin(2*t) + 0.3 + 0.7*bdy.normal_x, np.cos(t - 0.2) + 0.7*bdy.normal_y])
    mu = np.concatenate([np.cos(3*t + 0.4), 0.5*np.sin(2*t) + 0.1])
    # the double-layer density must carry no net flux: the velocity jumps across the curve by mu, so outside the flux is
    # int mu.n ds, which stokeslets inside the body cannot produce (in the solver mu is the jump of two divergence-free
    # fields and has none)
    n = bdy.N
    flux = np.sum((mu[:n]*bdy.normal_x + mu[n:]*bdy.normal_y)*bdy.weights)/np.sum(bdy.weights)
    mu = mu - flux*np.concatenate([bdy.normal_x, bdy.normal_y])
    q = Q.Stokes_QFS(bdy, interior, True, True, tol=1e-13)
    src = q([sig, mu]).reshape(2, -1)
    tt = np.linspace(0, 2*np.pi, 41, endpoint=False)
    z = (0.45 if interior else 1.9)*np.exp(1j*tt)*(1 + 0.1*np.cos(2*tt))
    K = q.kernel
    w2 = np.tile(q.source.weights, 2)
    val = K.slp(z.real, z.imag, q.source.x, q.source.y).dot(src.ravel()*w2)
    up = 4
    n = bdy.N
    fine = GSB(c=fourier_resample(bdy.c, up*n))
    rs = lambda a: fourier_resample(a, up*n).real
    sf = np.concatenate([rs(sig[:n]), rs(sig[n:])])
    mf = np.concatenate([rs(mu[:n]), rs(mu[n:])])
    wf = np.tile(fine.weights, 2)
    ref = K.slp(z.real, z.imag, fine.x, fine.y).dot(sf*wf) + \
        K.dlp(z.real, z.imag, fine.x, fine.y, fine.normal_x, fine.normal_y).dot(mf*wf)
    assert np.abs(val - ref).max() < 5e-12*max(1.0, np.abs(ref).max()), np.abs(val - ref).max()
