"""Multi-GPU check + timing of the slab-decomposed grid solve.  Launch:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
       scripts/run_slab_multi.py --grid 8192 [--check]
Every rank holds its slab of a seeded random forcing; --check gathers the result on rank 0 and compares with NumPy
(small n only).  Timing: CUDA events, max over ranks."""
import argparse, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch, torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument('--grid', '--n', dest='n', type=int, default=4096)
ap.add_argument('--steps', type=int, default=10)
ap.add_argument('--check', action='store_true')
ap.add_argument('--chunks', type=int, default=1)
ap.add_argument('--profile', action='store_true', help='per-kernel CUDA-event times of one solve (library profiler)')
a = ap.parse_args()
rank, world, lr = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
from ipde_b200.distributed import SlabGridSolver
import ipde_b200.device as D
n = a.n
xh = 2.4/n
S = SlabGridSolver(n, n, xh, xh, 1, 2.0, chunks=a.chunks)
L = S.L
if a.check:
    f = np.random.default_rng(0).standard_normal((n, n))
    fl = D.to_dev(f[L.row0:L.row0 + L.rows])
else:
    g = torch.Generator(device='cuda'); g.manual_seed(rank)
    fl = torch.randn((L.rows, n), dtype=torch.float64, device='cuda', generator=g)
out = S.solve(fl)
torch.cuda.synchronize()
if a.check:
    kx = np.fft.fftfreq(n, xh/(2*np.pi))[:, None]; ky = np.fft.fftfreq(n, xh/(2*np.pi))[None, :]
    ref = np.fft.ifft2(np.fft.fft2(f)/(4.0 + kx*kx + ky*ky)).real[L.row0:L.row0 + L.rows]
    err = float(np.abs(out.cpu().numpy() - ref).max()/np.abs(ref).max())
    errs = [None]*world
    if world > 1:
        dist.all_gather_object(errs, err)
    else:
        errs = [err]
    if rank == 0:
        print('slab solve n=%d world=%d: max rel err over ranks %.2e' % (n, world, max(errs)))
        assert max(errs) < 1e-13
for _ in range(3):
    S.solve(fl, out=out)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    S.solve(fl, out=out)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1)/a.steps], dtype=torch.float64, device='cuda')
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t[0])
    print(json.dumps({'slab_grid_solve': {'grid': [n, n], 'n_gpus': world, 'chunks': S.K, 'ms_per_solve': ms,
                                           'grid_point_solves_per_s': n*n/(ms*1e-3),
                                           'alltoall_bytes_per_rank_per_solve': S.alltoall_bytes_per_rank,
                                           'alg_GBps_64B_per_point': 64.0*n*n/(ms*1e-3)/1e9}}))
if a.profile and rank == 0:
    import ctypes
    from ipde_b200._lib import lib
    lib().ipde_profile_start(D.stream())
    S.solve(fl, out=out)
    buf = ctypes.create_string_buffer(1 << 16)
    lib().ipde_profile_stop(buf, len(buf))
    for line in buf.value.decode().strip().splitlines():
        k, nl, ms, by = line.split()
        print('%-24s n=%3s %8.3f ms  %8.1f GB/s' % (k, nl, float(ms), float(by)/(float(ms)*1e-3)/1e9 if float(ms) > 0 else 0))
if world > 1:
    dist.destroy_process_group()
