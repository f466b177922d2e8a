"""Multi-GPU check of boundary sharding: 4-boundary modified-Helmholtz solve with the boundaries distributed over the
ranks must equal the unsharded solve bit for bit.  Launch with torchrun (see scripts/run_slab_multi.py)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, torch, torch.distributed as dist
rank, world, lr = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
from problems import four_bodies, manufactured
from ipde_b200.solvers.multi_boundary.modified_helmholtz import ModifiedHelmholtzSolver
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
backend = sys.argv[2] if len(sys.argv) > 2 else 'ewald'
ebdyc = four_bodies(nb=nb, M=8)
f, ua, _ = manufactured(ebdyc, 5.0, kk=10*np.pi/7)
kw = dict(tol=1e-13, restart=60, maxiter=5)
sharded = ModifiedHelmholtzSolver(ebdyc, 5.0, grid_backend=backend, shard_boundaries=True)
us = np.asarray(sharded(f, **kw))
t0 = time.perf_counter()
for _ in range(5):
    sharded(f, **kw)
torch.cuda.synchronize()
ts = (time.perf_counter() - t0)/5
full = ModifiedHelmholtzSolver(ebdyc, 5.0, grid_backend=backend)
uf = np.asarray(full(f, **kw))
t0 = time.perf_counter()
for _ in range(5):
    full(f, **kw)
torch.cuda.synchronize()
tf = (time.perf_counter() - t0)/5
diff = float(np.abs(us - uf).max())
gathered = [None]*world
if world > 1:
    dist.all_gather_object(gathered, diff)
else:
    gathered = [diff]
if rank == 0:
    print('boundary sharding: world %d, owners %s, iterations %s, max |sharded - unsharded| over ranks %.2e, '
          'ms per solve sharded %.2f vs unsharded %.2f' % (world, sharded.owner, sharded.iteration_counts,
                                                            max(gathered), ts*1e3, tf*1e3))
    assert max(gathered) == 0.0

# ---- the same for the Stokes solver (three bodies): python ... run_shard_multi.py 100 ewald stokes
if len(sys.argv) > 3 and sys.argv[3] == 'stokes':
    from problems import three_bodies, stokes_manufactured
    from ipde_b200.solvers.multi_boundary.stokes import StokesSolver
    eb3 = three_bodies(nb=nb, M=8, bump=True)
    fu, fv, _, _, _, _ = stokes_manufactured(eb3)
    kws = dict(tol=1e-13, restart=80, maxiter=4)
    sh = StokesSolver(eb3, grid_backend=backend, shard_boundaries=True)
    fl = StokesSolver(eb3, grid_backend=backend)
    rs, rf = sh(fu, fv, **kws), fl(fu, fv, **kws)
    d = float(max(np.abs(np.asarray(a) - np.asarray(b)).max() for a, b in zip(rs, rf)))
    g2 = [None]*world
    if world > 1:
        dist.all_gather_object(g2, d)
    else:
        g2 = [d]
    if rank == 0:
        print('Stokes boundary sharding: world %d, owners %s, iterations %s / %s, max |sharded - unsharded| %.2e' %
              (world, sh.owner, sh.iteration_counts, fl.iteration_counts, max(g2)))
        assert max(g2) == 0.0
if world > 1:
    dist.destroy_process_group()
