"""Timing of the Stokes inhomogeneous solve on the geometry of
examples/interior_stokes_with_multi.py at production size:  python scripts/bench_stokes.py --grid 1024
Device-resident solves (CUDA events) and host-API solves (wall clock), plus the error against the analytic
solution after a homogeneous Dirichlet correction by stokeslets (velocity; checks the whole chain end to end)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, torch

ap = argparse.ArgumentParser()
ap.add_argument('--grid', type=int, default=1024)
ap.add_argument('--M', type=int, default=16)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--check', action='store_true')
ap.add_argument('--profile', action='store_true', help='per-kernel CUDA-event times of one solve')
a = ap.parse_args()
from problems import stokes_star, stokes_manufactured
from ipde_b200.solvers.multi_boundary.stokes import StokesSolver
from ipde_b200._lib import lib
import ipde_b200.device as D

t0 = time.perf_counter()
nb = 2*a.grid
ebdyc = stokes_star(nb=nb, M=a.M, ng=a.grid)
fu, fv, ua, va, pa, (u_f, v_f, p_f) = stokes_manufactured(ebdyc)
t_geo = time.perf_counter() - t0
t0 = time.perf_counter()
solver = StokesSolver(ebdyc, grid_backend='ewald')
torch.cuda.synchronize()
t_setup = time.perf_counter() - t0
kw = dict(tol=1e-12, restart=60, maxiter=4)
fud, fvd = D.to_dev(np.asarray(fu)), D.to_dev(np.asarray(fv))
for _ in range(2):
    res = solver.solve_device(fud, fvd, **kw)
torch.cuda.synchronize()
n0 = lib().ipde_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    res = solver.solve_device(fud, fvd, **kw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/a.steps
launches = (lib().ipde_launch_count() - n0)//a.steps
t0 = time.perf_counter()
for _ in range(a.steps):
    u, v, p = solver(fu, fv, **kw)
torch.cuda.synchronize()
ms_host = (time.perf_counter() - t0)*1e3/a.steps
out = {'stokes_solve': {'grid': [a.grid, a.grid], 'n_boundary': nb, 'M': a.M, 'dof': int(ebdyc.dof),
                        'evaluation': solver.evaluation, 'gmres_iterations': solver.iteration_counts,
                        'ms_per_solve_device': ms, 'ms_per_solve_host_api': ms_host,
                        'grid_point_solves_per_s': a.grid*a.grid/(ms*1e-3), 'kernel_launches_per_solve': int(launches),
                        'setup_s': {'geometry_host': t_geo, 'solver': t_setup}}}
if a.check:
    from ipde_b200.qfs import Stokes_QFS
    bdy = ebdyc[0].bdy
    bu = solver.helpers[0].get_boundary_values(u[0])
    bv = solver.helpers[0].get_boundary_values(v[0])
    qb = Stokes_QFS(bdy, True, True, True, tol=1e-13)
    sig = qb.u2s(np.concatenate([u_f(bdy.x, bdy.y) - bu, v_f(bdy.x, bdy.y) - bv])).reshape(2, -1)
    pts = ebdyc.grid_and_radial_pts
    # a sample of the targets keeps the all-pairs correction cheap
    idx = np.random.default_rng(0).choice(pts.N, size=min(pts.N, 20000), replace=False)
    from ipde_b200.geometry import PointSet
    sub = PointSet(x=pts.x[idx], y=pts.y[idx])
    cu, cv, cp = solver.Layer_Apply(qb.source, sub, sig)
    err = max(np.abs(np.asarray(u)[idx] + cu - np.asarray(ua)[idx]).max(),
              np.abs(np.asarray(v)[idx] + cv - np.asarray(va)[idx]).max())
    out['stokes_solve']['max_velocity_error_vs_analytic'] = float(err)
if a.profile:
    import ctypes
    lib().ipde_profile_start(D.stream())
    solver.solve_device(fud, fvd, **kw)
    buf = ctypes.create_string_buffer(1 << 16)
    lib().ipde_profile_stop(buf, len(buf))
    prof = {}
    for line in buf.value.decode().strip().splitlines()[:14]:
        k, nl, t, by = line.split()
        prof[k] = {'launches': int(nl), 'ms': float(t)}
    out['stokes_solve']['kernels_ms_per_solve'] = prof
print(json.dumps(out))
