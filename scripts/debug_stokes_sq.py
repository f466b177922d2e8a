import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from ipde_b200.geometry import GSB, star
from ipde_b200 import qfs as Q
import ipde_b200.qfs_structured as S
n = int(sys.argv[1])
bdy = GSB(c=star(n, a=0.1, f=5))
orig = S._significant
def spy(X, Bx, Ks, att, tol, far):
    sig = orig(X, Bx, Ks, att, tol, far)
    ncol = X.shape[3]; h_in = (ncol-1)//2
    p = torch.arange(-Bx, Bx+1, device=X.device); kp = torch.arange(-h_in, h_in+1, device=X.device)
    pa = p.abs()[:, None].expand(sig.shape)
    for thr in (400, 300, 200, 150):
        idx = torch.nonzero(sig & (pa >= thr))
        print('  |p|>=%d significant entries: %d' % (thr, idx.shape[0]), [(int(p[i]), int(kp[j])) for i, j in idx[:5]], [(int(p[i]), int(kp[j])) for i, j in idx[-3:]])
    return sig
S._significant = spy
t0 = time.time()
q = S.StructuredQFS(bdy, False, True, True, Q.StokesKernel(), tol=1e-12)
print('setup', time.time() - t0, (q.op.KL, q.op.BH), (q.op_u2s.KL, q.op_u2s.BH))
