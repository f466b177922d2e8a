"""prints the achieved device-vs-oracle errors of the Stokes parity cases for several GMRES tolerances"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from problems import stokes_star, stokes_manufactured, three_bodies
from ipde_b200.solvers.multi_boundary.stokes import StokesSolver
from oracle.solvers import OracleVectorSolver
def run(name, ebdyc, tol, **kw):
    fu, fv, ua, va, pa, _ = stokes_manufactured(ebdyc)
    solver = StokesSolver(ebdyc, **kw)
    gm = dict(tol=tol, restart=80, maxiter=4)
    u, v, p = solver(fu, fv, **gm)
    pairs = [(h.interface_qfs_g, h.interface_qfs_r) for h in solver.helpers]
    O = OracleVectorSolver(ebdyc, kw.get('solver_type', 'spectral'), pairs)
    uo, vo, po = O(fu, fv, **gm)
    scale = max(np.abs(uo).max(), np.abs(vo).max())
    eu = np.abs(np.asarray(u) - uo).max()/scale; ev = np.abs(np.asarray(v) - vo).max()/scale
    dp = np.asarray(p) - po
    pna = ebdyc.phys_inds[ebdyc.phys_not_in_annulus]
    ep = np.abs(dp[pna]).max()/np.abs(po).max()
    n0 = ebdyc.grid_phys.N
    epr = []
    for b in range(len(ebdyc.ebdys)):
        blk = dp[n0 + sum(ebdyc.radial_dof_list[:b]):n0 + sum(ebdyc.radial_dof_list[:b + 1])]
        epr.append(np.abs(blk - blk.mean()).max()/np.abs(po).max())
    print('%-34s tol %.0e: u %.1e v %.1e p(grid) %.1e p(annuli, up to const) %s  its %s / %s' % (name, tol, eu, ev, ep, ['%.1e' % x for x in epr], solver.iteration_counts, O.iteration_counts), flush=True)
for tol in (1e-13, 1e-15):
    run('single spectral', stokes_star(nb=200, M=8), tol)
    run('single fourth', stokes_star(nb=200, M=8), tol, solver_type='fourth')
    run('single ewald', stokes_star(nb=200, M=8), tol, grid_backend='ewald')
    run('three bodies direct', three_bodies(nb=100, M=8, bump=True), tol)
    run('three bodies ewald', three_bodies(nb=100, M=8, bump=True), tol, grid_backend='ewald')
