"""Flakiness hunt: same single-boundary problem solved repeatedly, grid/radial error after an MFS Dirichlet correction."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, torch
from problems import single_star, manufactured
from ipde_b200.solvers.multi_boundary.modified_helmholtz import ModifiedHelmholtzSolver
from ipde_b200.qfs import Modified_Helmholtz_QFS
import ipde_b200.device as D
k, kk = 2.0, np.pi/3
for ng in (480, 240, 480, 240):
    junk = torch.full((50_000_000,), float('nan'), dtype=torch.float64, device='cuda'); del junk     # poison freed memory
    ebdyc = single_star(nb=400, M=12, ng=ng)
    f, ua, sol = manufactured(ebdyc, k, kk=kk)
    solver = ModifiedHelmholtzSolver(ebdyc, k)
    bdy = ebdyc[0].bdy
    qb = Modified_Helmholtz_QFS(bdy, True, True, False, k, tol=1e-12, b2c=False)
    pts = ebdyc.grid_and_radial_pts
    nphys = int(ebdyc.phys.sum())
    for rep in range(3):
        junk = torch.full((20_000_000,), float('nan'), dtype=torch.float64, device='cuda'); del junk
        ue = solver(f, tol=1e-13, restart=60, maxiter=5)
        g = sol(bdy.x, bdy.y) - np.asarray(solver.get_boundary_values(ue))
        sig = qb.u2s(g)
        corr = D.layer_apply_scalar(1, k, D.to_dev(qb.source.x), D.to_dev(qb.source.y), D.to_dev(sig*qb.source.weights),
                                    D.to_dev(pts.x), D.to_dev(pts.y)).cpu().numpy()
        err = np.abs(np.asarray(ue) + corr - np.asarray(ua))
        bad = np.nonzero(err[:nphys] > 1e-6)[0]
        print('ng', ng, 'rep', rep, 'grid err %.2e radial err %.2e' % (np.nanmax(err[:nphys]), np.nanmax(err[nphys:])),
              'nan', int(np.isnan(err).sum()), 'bad grid pts', bad.size,
              ('first bad xy (%.3f, %.3f)' % (pts.x[bad[0]], pts.y[bad[0]])) if bad.size else '')
