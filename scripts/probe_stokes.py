"""Scaling probe, Stokes: BASELINE configs[3] geometry at grid size N; setup phases, solve stages, error vs analytic."""
import sys, os, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
M = int(os.environ.get('PROBE_M', '20'))
def tic(name, t0):
    print('%-28s %8.3f s' % (name, time.time() - t0), flush=True)
import torch
from problems import stokes_star_box, stokes_manufactured
t0 = time.time()
ebdyc = stokes_star_box(N, M=M)
tic('ebdy + register_grid', t0)
print('n_b', [e.bdy.N for e in ebdyc], 'dof', ebdyc.dof)
from ipde_b200.solvers.multi_boundary.stokes import StokesSolver
t0 = time.time()
fu, fv, ua, va, pa, _ = stokes_manufactured(ebdyc)
tic('forcing', t0)
t0 = time.time()
solver = StokesSolver(ebdyc)
torch.cuda.synchronize()
tic('solver ctor', t0)
for h in solver.helpers:
    q, r = h.interface_qfs_g, h.interface_qfs_r
    print(' qfs', type(q).__name__, 'N', q.N, 'M', q.M, getattr(q, 'J', None),
          [(o.KL, o.BH) for o in (q.op, r.op, r.op_u2s)] if getattr(q, 'structured', False) else '',
          'curve ops', None if h.curve_r is None else (h.curve_r.op.KL, h.curve_r.op.BH))
print('evaluation', solver.evaluation, 'mem GB', torch.cuda.max_memory_allocated()/1e9)
gm = dict(tol=1e-12, restart=50, maxiter=4)
t0 = time.time()
ue, ve, pe = solver(fu, fv, **gm)
tic('first solve (host API)', t0)
print('gmres its', solver.iteration_counts)
# the inhomogeneous solution differs from the analytic one by a homogeneous Stokes field; report the raw difference
# only as a sanity figure (not an error)
from ipde_b200.device import to_dev
fu_d, fv_d = to_dev(np.asarray(fu)), to_dev(np.asarray(fv))
for _ in range(2):
    res = solver.solve_device(fu_d, fv_d, **gm)
torch.cuda.synchronize()
t0 = time.time()
for _ in range(5):
    res = solver.solve_device(fu_d, fv_d, **gm)
torch.cuda.synchronize()
dt = (time.time() - t0)/5
print('solve_device %.3f ms  -> %.3e grid-point-solves/s' % (dt*1e3, N*N/dt))
print('peak mem GB', torch.cuda.max_memory_allocated()/1e9)
import ctypes
from ipde_b200._lib import lib
from ipde_b200.device import stream
lib().ipde_profile_start(stream())
solver.solve_device(fu_d, fv_d, **gm)
buf = ctypes.create_string_buffer(1 << 16)
lib().ipde_profile_stop(buf, len(buf))
print('kernel launches ms bytes GB/s')
for line in buf.value.decode().strip().splitlines():
    pr = line.split()
    if len(pr) == 4 and float(pr[2]) > 0.1:
        print('  %-28s %5d %9.3f %12.4g %8.0f' % (pr[0], int(pr[1]), float(pr[2]), float(pr[3]), float(pr[3])/float(pr[2])/1e6))
json.dump({'N': N, 'solve_ms': dt*1e3, 'gmres': solver.iteration_counts}, open(os.path.join(ROOT, 'gpurun_out', 'probe_stokes_%d.json' % N), 'w'))
