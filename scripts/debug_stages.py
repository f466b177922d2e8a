"""Stage-by-stage comparison of the device scalar solve with the CPU oracle (debug aid, run on a GPU box)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np
from problems import single_star, three_bodies, manufactured
from ipde_b200.solvers.multi_boundary.modified_helmholtz import ModifiedHelmholtzSolver
from ipde_b200.solvers.multi_boundary.poisson import PoissonSolver
from oracle.solvers import OracleScalarSolver
import ipde_b200.device as D

def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max()/max(np.abs(b).max(), 1e-300)

kind = sys.argv[1] if len(sys.argv) > 1 else 'mh'
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 200
if kind == 'mh':
    ebdyc = single_star(nb=nb, M=8)
    f, ua, _ = manufactured(ebdyc, 2.0)
    solver = ModifiedHelmholtzSolver(ebdyc, 2.0)
    okind, k = 'modified_helmholtz', 2.0
else:
    ebdyc = single_star(nb=nb, M=8, bump=True)
    f, ua, _ = manufactured(ebdyc, 0.0)
    solver = PoissonSolver(ebdyc)
    okind, k = 'poisson', 0.0
print('grid', ebdyc.grid.Nx, 'layout', solver.dev.grid.n1, solver.dev.grid.n2)
kw = dict(tol=1e-13, restart=40, maxiter=5)
ue = np.asarray(solver(f, **kw))
pairs = [(h.interface_qfs_g, h.interface_qfs_r) for h in solver.helpers]
O = OracleScalarSolver(ebdyc, okind, k, 'spectral', pairs)
uo = O(f, **kw)
Dv = solver.dev
print('uc (grid solve)      ', rel(Dv.uc.cpu().numpy(), O.uc_grid_only))
for name, a, b in zip(['bv', 'bx', 'by'], solver._interface_values, O.interface_values):
    print('interface', name, '        ', rel(a.cpu().numpy(), b))
h, ho = solver.helpers[0], O.helpers[0]
print('iterations', solver.iteration_counts, O.iteration_counts)
print('sigma_g              ', rel(h.sigma_g.cpu().numpy(), ho.sigma_g), np.abs(ho.sigma_g).max())
print('sigma_r              ', rel(h.sigma_r.cpu().numpy(), ho.sigma_r))
print('tau                  ', rel(h._tau.cpu().numpy(), np.concatenate([ho.slp, ho.dlp])))
print('ur (after correct)   ', rel(h.ur.cpu().numpy(), ho.ur))
n = ebdyc.grid_phys.N
print('final grid part      ', rel(ue[:n], uo[:n]))
print('final radial part    ', rel(ue[n:], uo[n:]))
pna = ebdyc.phys_inds[ebdyc.phys_not_in_annulus]
ia = ebdyc[0].ia_inds
print('  pna points         ', rel(ue[pna], uo[pna]))
print('  in-annulus points  ', rel(ue[ia], uo[ia]))
