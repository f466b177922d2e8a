import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import ipde_b200.device as D
from ipde_b200.geometry import GSB, star
from ipde_b200 import qfs as Q
import ipde_b200.qfs_structured as S
from ipde_b200.solvers.internals.scalar import device_qfs
n = int(sys.argv[1])
bdy = GSB(c=star(n, a=0.1, f=5))
res = {}
for cap in ('0', '256'):
    os.environ['IPDE_QFS_BH_CAP'] = cap
    q = S.StructuredQFS(bdy, False, True, True, Q.StokesKernel(), tol=1e-12)
    print('cap', cap, (q.op.KL, q.op.BH), (q.op_u2s.KL, q.op_u2s.BH))
    dq = device_qfs(q)
    t = bdy.t
    rng = np.random.default_rng(1)
    tau = np.concatenate([np.exp(np.cos(t))*np.sin(2*t) + 1e-9*rng.standard_normal(n), np.cos(t), np.cos(3*t + 0.3), np.sin(2*t)])
    sig = dq(D.to_dev(tau))
    ub = np.concatenate([np.cos(3*t), np.sin(5*t)])
    su = dq.u2s(D.to_dev(ub))
    # velocity of the sources at targets one node spacing outside the curve (evaluation side: exterior)
    th = np.linspace(0, 2*np.pi, 200, endpoint=False)
    z = (1 + 0.1*np.cos(5*th))*np.exp(1j*th)*(1 + 2*np.pi/n)
    tx, ty = D.to_dev(z.real.copy()), D.to_dev(z.imag.copy())
    for name, s in (('b2c', sig), ('u2s', su)):
        s2 = s.view(2, -1)
        out = D.layer_apply_stokes(dq.sx, dq.sy, s2[0].contiguous(), s2[1].contiguous(), tx, ty).cpu().numpy()
        res[(cap, name)] = out
for name in ('b2c', 'u2s'):
    a, b = res[('0', name)], res[('256', name)]
    print(name, 'field difference capped vs full band: %.2e (scale %.2e)' % (np.abs(a - b).max(), np.abs(a).max()))
