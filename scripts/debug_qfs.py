import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, scipy.linalg
from problems import single_star, manufactured
from ipde_b200.qfs import Modified_Helmholtz_QFS
import ipde_b200.device as D
ebdyc = single_star(nb=200, M=8)
e = ebdyc[0]
for interior in [True, False]:
    q = Modified_Helmholtz_QFS(e.interface, interior, True, True, 2.0, tol=1e-12)
    rng = np.random.default_rng(0)
    t = e.bdy.t
    tau = np.concatenate([np.exp(np.sin(3*t)), np.cos(2*t)*np.exp(np.cos(t))])
    v = q.C.dot(tau)
    ref = scipy.linalg.lu_solve(q.E_lu, v)
    lu, piv = q.E_lu
    L = np.tril(lu, -1) + np.eye(lu.shape[0]); U = np.triu(lu)
    perm = np.arange(lu.shape[0])
    for i, p in enumerate(piv):
        perm[i], perm[p] = perm[p], perm[i]
    E = (L @ U)
    dl = D.DenseLU(q.E_lu)
    sol = dl.solve(D.to_dev(v)).cpu().numpy()
    # residuals in the permuted system  L U x = P v
    print('interior', interior, 'M', q.M, '|sigma| ref %.2e gpu %.2e' % (np.abs(ref).max(), np.abs(sol).max()))
    print('  residual ref %.2e  gpu %.2e  (|v| %.2e)' % (np.abs(E.dot(ref) - v[perm]).max(), np.abs(E.dot(sol) - v[perm]).max(), np.abs(v).max()))
    # forward step only
    y_ref = scipy.linalg.solve_triangular(L, v[perm], lower=True, unit_diagonal=True)
    print('  cond-ish: |y| %.2e, diag U min %.2e max %.2e' % (np.abs(y_ref).max(), np.abs(np.diag(U)).min(), np.abs(np.diag(U)).max()))
