"""Shared small problem setups for the parity tests (seeded / deterministic geometry)."""
import numpy as np
from ipde_b200.geometry import GSB, Grid, star, squished_circle
from ipde_b200.embedded_boundary import EmbeddedBoundary
from ipde_b200.ebdy_collection import EmbeddedBoundaryCollection
from ipde_b200.embedded_function import EmbeddedFunction
from ipde_b200.slepian import SlepianMollifier


def single_star(nb=200, M=8, a=0.1, f=5, ng=None, bump=False, slepian_r=None):
    """examples/interior_modified_helmholtz.py geometry: star in [-1.2, 1.2]^2"""
    bdy = GSB(c=star(nb, a=a, f=f))
    bh = bdy.dt*bdy.speed.min()
    if ng is None:
        ng = 2*int(0.5*2.4//bh)
    grid = Grid([-1.2, 1.2], ng, [-1.2, 1.2], ng)
    MOL = SlepianMollifier(2*M if slepian_r is None else slepian_r)
    ebdy = EmbeddedBoundary(bdy, True, M, bh, heaviside=MOL.step)
    ebdyc = EmbeddedBoundaryCollection([ebdy])
    ebdyc.register_grid(grid)
    if bump:
        ebdyc.ready_bump(MOL.bump, (1.2 - ebdy.radial_width, 1.2 - ebdy.radial_width), ebdy.radial_width)
    return ebdyc


def three_bodies(nb=100, M=8, bump=False):
    """scaled-down examples/multi_modified_helmholtz.py: interior star + two exterior bodies"""
    bdy1 = GSB(c=star(4*nb, r=1.0, a=0.05, f=7))
    bdy2 = GSB(c=squished_circle(nb, x=-0.4, y=-0.25, r=0.25, b=0.6, rot=-np.pi/4))
    bdy3 = GSB(c=star(nb, x=0.35, y=0.2, r=0.22, a=0.1, f=3))
    bdys = [bdy1, bdy2, bdy3]
    bh = min(b.dt*b.speed.min() for b in bdys)
    MOL = SlepianMollifier(2*M)
    ebdys = [EmbeddedBoundary(b, b is bdy1, M, bh, heaviside=MOL.step) for b in bdys]
    ebdyc = EmbeddedBoundaryCollection(ebdys)
    ebdyc.generate_grid(force_square=True)
    if bump:
        ebdyc.ready_bump(MOL.bump)
    return ebdyc


def manufactured(ebdyc, helmholtz_k, kk=8*np.pi/3):
    sol = lambda x, y: np.exp(np.sin(kk*x))*np.sin(kk*y)
    lap = lambda x, y: kk**2*np.exp(np.sin(kk*x))*np.sin(kk*y)*(np.cos(kk*x)**2 - np.sin(kk*x) - 1.0)
    force = lambda x, y: helmholtz_k**2*sol(x, y) - lap(x, y)
    f = EmbeddedFunction(ebdyc, function=force if helmholtz_k != 0 else lap)
    ua = EmbeddedFunction(ebdyc, function=sol)
    return f, ua, sol


def stokes_star(nb=200, M=8, ng=None):
    """examples/interior_stokes_with_multi.py geometry: star in [-pi/2, pi/2]^2, radial spacing 0.75*grid spacing"""
    bdy = GSB(c=star(nb, a=0.1, f=5))
    if ng is None:
        ng = 2*int(nb/4)
    L = np.pi/2
    grid = Grid([-L, L], ng, [-L, L], ng)
    MOL = SlepianMollifier(2*M)
    ebdy = EmbeddedBoundary(bdy, True, M, grid.xh*0.75, heaviside=MOL.step)
    ebdyc = EmbeddedBoundaryCollection([ebdy])
    ebdyc.register_grid(grid)
    ebdyc.ready_bump(MOL.bump, (L - ebdy.radial_width, L - ebdy.radial_width), ebdy.radial_width)
    return ebdyc


def stokes_manufactured(ebdyc):
    """examples/interior_stokes_with_multi.py:58-85"""
    a, b, p_a, p_b = 3.0, 2.0, 2.0, 1.0
    sin, cos = np.sin, np.cos
    esin = lambda x: np.exp(sin(x))
    psix = lambda x, y: esin(a*x)*cos(b*y)
    psiy = lambda x, y: esin(a*x)*sin(b*y)
    u_f = lambda x, y: psix(x, y)
    v_f = lambda x, y: -a/b*cos(a*x)*psiy(x, y)
    p_f = lambda x, y: cos(p_a*x) + esin(p_b*y)
    fu_f = lambda x, y: (a**2*(sin(a*x) - cos(a*x)**2) + b**2)*psix(x, y) - p_a*sin(p_a*x)
    fv_f = lambda x, y: -a*b*cos(a*x)*psiy(x, y)*(1 + (a/b)**2*sin(a*x)*(3 + sin(a*x))) + p_b*cos(p_b*y)*esin(p_b*y)
    E = lambda f: EmbeddedFunction(ebdyc, function=f)
    return E(fu_f), E(fv_f), E(u_f), E(v_f), E(p_f), (u_f, v_f, p_f)


def four_bodies(nb=100, M=8, bump=False):
    """BASELINE configs[2]: examples/multi_modified_helmholtz.py:34-36 (interior star + exterior squished circle +
    exterior star), scaled to a unit interior star, plus the 4th exterior body BASELINE asks for (declared in
    SURVEY.md 8(d).3: a 4-armed star)."""
    bdy1 = GSB(c=star(4*nb, r=1.0, a=0.05, f=7))
    bdy2 = GSB(c=squished_circle(nb, x=-0.4, y=-0.25, r=0.25, b=0.6, rot=-np.pi/4))
    bdy3 = GSB(c=star(nb, x=0.35, y=0.2, r=0.22, a=0.1, f=3))
    bdy4 = GSB(c=star(nb, x=-0.17, y=0.5, r=0.2, a=0.15, f=4))
    bdys = [bdy1, bdy2, bdy3, bdy4]
    bh = min(b.dt*b.speed.min() for b in bdys)
    MOL = SlepianMollifier(2*M)
    ebdys = [EmbeddedBoundary(b, b is bdy1, M, bh, heaviside=MOL.step) for b in bdys]
    ebdyc = EmbeddedBoundaryCollection(ebdys)
    ebdyc.generate_grid(force_square=True)
    if bump:
        ebdyc.ready_bump(MOL.bump)
    return ebdyc


def four_bodies_box(N, M=20, nb_outer=None, nb_small=None, bump=False, three=False):
    """
    BASELINE configs[2] / [4] at full scale: the geometry of examples/multi_modified_helmholtz.py:34-36 on its own box
    [-3.5, 3.5]^2 (interior star r = 3, a = 0.05, f = 7; exterior squished circle at (-1.2, -0.7); exterior star
    a = 0.2, f = 3 at (1, 0.5)) plus the 4th exterior body declared in SURVEY.md 8(d).3 (star x = -0.5, y = 1.5,
    r = 0.6, a = 0.15, f = 4) on an N x N grid.  Boundary node counts default to 3N (outer), 3N/4 (unit bodies) and
    3N/8 (4th body), i.e.
    node spacing ~ grid spacing, with FFT-friendly lengths; the radial width is M grid cells.
    """
    L = 3.5
    grid = Grid([-L, L], N, [-L, L], N)
    h = grid.xh
    n1 = 3*N if nb_outer is None else nb_outer
    n2 = (3*N)//4 if nb_small is None else nb_small
    bdy1 = GSB(c=star(n1, r=3.0, a=0.05, f=7))
    bdy2 = GSB(c=squished_circle(n2, x=-1.2, y=-0.7, r=1.0, b=0.6, rot=-np.pi/4))
    bdy3 = GSB(c=star(n2, x=1.0, y=0.5, r=1.0, a=0.2, f=3))
    bdy4 = GSB(c=star(max(n2//2, 16), x=-0.5, y=1.5, r=0.6, a=0.15, f=4))
    bdys = [bdy1, bdy2, bdy3] if three else [bdy1, bdy2, bdy3, bdy4]
    MOL = SlepianMollifier(2*M)
    ebdys = [EmbeddedBoundary(b, b is bdy1, M, h, heaviside=MOL.step) for b in bdys]
    ebdyc = EmbeddedBoundaryCollection(ebdys)
    ebdyc.register_grid(grid)
    if bump:
        ebdyc.ready_bump(MOL.bump, (L - ebdys[0].radial_width, L - ebdys[0].radial_width), ebdys[0].radial_width)
    return ebdyc


def stokes_star_box(N, M=20, nb=None):
    """BASELINE configs[3] at full scale: examples/interior_stokes_with_multi.py geometry (star a=0.1, f=5 in
    [-pi/2, pi/2]^2, radial spacing 0.75 * grid spacing) on an N x N grid; nb defaults to 3N (node spacing ~ the
    radial spacing, FFT-friendly)."""
    return stokes_star(nb=3*N if nb is None else nb, M=M, ng=N)


def interior_star_box(ng, M=16, helmholtz_k=2.0):
    """BASELINE configs[1]: examples/interior_modified_helmholtz.py scaled up -- one star (a=0.1, f=5) in [-1.2, 1.2]^2
    on an ng x ng grid, boundary spacing matched to the grid spacing (nb = 2 ng), manufactured forcing of :49-52.
    Returns (ebdyc, f, nb)."""
    nb = 2*ng
    bdy = GSB(c=star(nb, a=0.1, f=5))
    bh = bdy.dt*bdy.speed.min()
    grid = Grid([-1.2, 1.2], ng, [-1.2, 1.2], ng)
    MOL = SlepianMollifier(2*M)
    ebdy = EmbeddedBoundary(bdy, True, M, bh, heaviside=MOL.step)
    ebdyc = EmbeddedBoundaryCollection([ebdy])
    ebdyc.register_grid(grid)
    kk = np.pi/3
    sol = lambda x, y: np.exp(np.sin(kk*x))*np.sin(kk*y)
    force = lambda x, y: helmholtz_k**2*sol(x, y) - \
        kk**2*np.exp(np.sin(kk*x))*np.sin(kk*y)*(np.cos(kk*x)**2 - np.sin(kk*x) - 1.0)
    return ebdyc, EmbeddedFunction(ebdyc, function=force), nb
