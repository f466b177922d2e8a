"""Exploration: inhomogeneous solve + DirichletCorrection against the analytic solution (prints errors)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np
from problems import single_star, three_bodies, four_bodies, manufactured
from ipde_b200.solvers.multi_boundary.modified_helmholtz import ModifiedHelmholtzSolver
from ipde_b200.solvers.multi_boundary.poisson import PoissonSolver
from ipde_b200.solvers.multi_boundary.homogeneous import DirichletCorrection
from ipde_b200.embedded_function import BoundaryFunction

def run(name, ebdyc, k, kk, poisson=False):
    f, ua, sol = manufactured(ebdyc, k, kk=kk)
    if poisson:
        solver = PoissonSolver(ebdyc)
    else:
        solver = ModifiedHelmholtzSolver(ebdyc, k)
    ue = solver(f, tol=1e-13, restart=60, maxiter=5)
    bc = BoundaryFunction(ebdyc, function=sol)
    for ev in ('direct', 'ewald'):
        corr = DirichletCorrection(solver, evaluation=ev)
        u = corr(ue, bc)
        err = np.abs(np.asarray(u) - np.asarray(ua)).max()
        print(name, ev, 'its', corr.iterations, 'err vs analytic %.3e' % err, 'before %.3e' % np.abs(np.asarray(ue) - np.asarray(ua)).max())

#run('single MH nb=400', single_star(nb=400, M=12, ng=480), 2.0, np.pi/3)
#run('three MH nb=200', three_bodies(nb=200, M=12), 5.0, 10*np.pi/7)
#run('four MH nb=200', four_bodies(nb=200, M=12), 5.0, 10*np.pi/7)

# ---- single-boundary diagnosis
from ipde_b200.qfs import Modified_Helmholtz_QFS
import ipde_b200.device as D
for nb, ng, M in ((400, 480, 12), (200, None, 8), (400, 240, 12)):
    ebdyc = single_star(nb=nb, M=M, ng=ng)
    k, kk = 2.0, np.pi/3
    f, ua, sol = manufactured(ebdyc, k, kk=kk)
    solver = ModifiedHelmholtzSolver(ebdyc, k)
    ue = solver(f, tol=1e-13, restart=60, maxiter=5)
    print('single nb', nb, 'grid', ebdyc.grid.Nx, 'its', solver.iteration_counts)
    bdy = ebdyc[0].bdy
    g = sol(bdy.x, bdy.y) - np.asarray(solver.get_boundary_values(ue))
    pts = ebdyc.grid_and_radial_pts
    nphys = int(ebdyc.phys.sum())
    for label, qb in (('slp+dlp tol 1e-13', Modified_Helmholtz_QFS(bdy, True, True, True, k, tol=1e-13)),
                      ('slp only b2c=False', Modified_Helmholtz_QFS(bdy, True, True, False, k, tol=1e-12, b2c=False))):
        sig = qb.u2s(g)
        corr = D.layer_apply_scalar(1, k, D.to_dev(qb.source.x), D.to_dev(qb.source.y), D.to_dev(sig*qb.source.weights),
                                    D.to_dev(pts.x), D.to_dev(pts.y)).cpu().numpy()
        err = np.abs(np.asarray(ue) + corr - np.asarray(ua))
        onb = D.layer_apply_scalar(1, k, D.to_dev(qb.source.x), D.to_dev(qb.source.y), D.to_dev(sig*qb.source.weights),
                                   D.to_dev(bdy.x), D.to_dev(bdy.y)).cpu().numpy()
        print('   ', label, 'grid err %.2e radial err %.2e  boundary residual %.2e  |g| %.2e |sig| %.2e' %
              (err[:nphys].max(), err[nphys:].max(), np.abs(onb - g).max(), np.abs(g).max(), np.abs(sig).max()))
    if nb == 400 and ng == 480:
        corr = DirichletCorrection(solver, evaluation='direct')
        t = D.to_dev(g)
        qd = corr._sources([t], D.empty((int(corr.soff[-1]),))).cpu().numpy()
        qh = corr.qfs[0].u2s(g)*corr.qfs[0].source.weights
        print('    class sources vs host u2s: %.2e (|q| %.2e)' % (np.abs(qd - qh).max(), np.abs(qh).max()))
        bc = BoundaryFunction(ebdyc, function=sol)
        print('    bc vs sol(bdy): %.2e' % np.abs(np.asarray(bc) - sol(bdy.x, bdy.y)).max())
        u = corr(ue, bc)
        err = np.abs(np.asarray(u) - np.asarray(ua))
        print('    class: grid err %.2e radial err %.2e' % (err[:nphys].max(), err[nphys:].max()))
