import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, torch
from problems import three_bodies, manufactured, stokes_manufactured
from ipde_b200.solvers.multi_boundary.poisson import PoissonSolver
from ipde_b200.solvers.multi_boundary.stokes import StokesSolver
import ipde_b200.device as D
ebdyc = three_bodies(nb=100, M=8, bump=True)
f, ua, _ = manufactured(ebdyc, 0.0)
kw = dict(tol=1e-13, restart=80, maxiter=4)
se = PoissonSolver(ebdyc, grid_backend='ewald')
ue = np.asarray(se(f, **kw))
sd = PoissonSolver(ebdyc, grid_backend='direct', AS_list=None)
ud = np.asarray(sd(f, **kw))
print('poisson ewald vs direct', np.abs(ue - ud).max()/np.abs(ud).max())
fu, fv, *_ = stokes_manufactured(ebdyc)
S1 = StokesSolver(ebdyc, grid_backend='ewald'); S2 = StokesSolver(ebdyc, grid_backend='direct')
a = S1(fu, fv, **kw); b = S2(fu, fv, **kw)
for k, nm in enumerate('uvp'):
    d = np.abs(np.asarray(a[k]) - np.asarray(b[k]))
    print('stokes', nm, d.max()/np.abs(np.asarray(b[k])).max(), 'grid part', d[:ebdyc.grid_phys.N].max(), 'radial part', d[ebdyc.grid_phys.N:].max())
# evaluator_w check per helper: own g sources -> own interface
for h in S1.helpers:
    q = h._weighted(h.sigma_g, h.d_qfs_g.sw)
    w1 = h.evaluator_w(q[0], q[1]).cpu().numpy()
    w2 = D.layer_apply_stokes(h.d_qfs_g.sx, h.d_qfs_g.sy, q[0], q[1], h.d_ix, h.d_iy).cpu().numpy()
    print('evaluator_w err', np.abs(w1 - w2).max(axis=1), 'scale', np.abs(w2).max(axis=1), 'box', h.evaluator_w.ev.px, h.evaluator_w.ev.L)
    qr = h._weighted(h.sigma_r, h.d_qfs_r.sw)
    r1 = h.evaluator_r(qr[0], qr[1]).cpu().numpy()
    r2 = D.layer_apply_stokes(h.d_qfs_r.sx, h.d_qfs_r.sy, qr[0], qr[1], h.d_rx, h.d_ry).cpu().numpy()
    print('evaluator_r err', np.abs(r1 - r2).max(axis=1), 'scale', np.abs(r2).max(axis=1))
# evaluator_g directly: all g-sources -> pna grid + interfaces vs the all-pairs kernel
Dv = S1.dev
sig = torch.cat([h._weighted(h.sigma_g, h.d_qfs_g.sw) for h in S1.helpers], dim=1).contiguous()
ref = D.layer_apply_stokes(S1.d_gsx, S1.d_gsy, sig[0].contiguous(), sig[1].contiguous(), Dv.pnai_x, Dv.pnai_y).cpu().numpy()
res = [D.zeros((ebdyc.dof,)) for _ in range(3)]
iface = S1.evaluator_g(sig[0].contiguous(), sig[1].contiguous(), grid_out=(res[0], res[1], res[2])).cpu().numpy()
pos = Dv.pna_pos.cpu().numpy()
for k, nm in enumerate('uvp'):
    g = res[k].cpu().numpy()[pos]
    print('evaluator_g', nm, 'grid err', np.abs(g - ref[k, :Dv.n_pna]).max(), 'iface err', np.abs(iface[k] - ref[k, Dv.n_pna:]).max(),
          'scale', np.abs(ref[k]).max(), 'mean grid diff', (g - ref[k, :Dv.n_pna]).mean())
