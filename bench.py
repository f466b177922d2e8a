#!/usr/bin/env python
"""
bench.py -- throughput of ipde's inhomogeneous-solve hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--n GRID]

Workload (BASELINE.json configs[1]): examples/interior_modified_helmholtz.py scaled up -- one star-shaped
interior boundary (a=0.1, f=5) in the box [-1.2, 1.2]^2, modified Helmholtz k=2, manufactured forcing,
FP64, spectral interface interpolation, annular GMRES tol 1e-12.  A "step" is one call of the solver on one
forcing (= Nx*Ny grid-point-solves).  metric = grid-point-solves/s.

  value   whole job, inputs already resident in HBM (solver.solve_device), CUDA-event timed
  e2e     the same through the public API solver(f) with host arrays: H2D of f and D2H of u inside the timing
  N > 1   `value`: every rank solves its own forcing on its own copy of the problem (independent solves are the
          unit that shards without a data-path collective: "scaling": "weak").  `slab_grid_solve`: ONE 16384^2
          periodic grid solve slab-decomposed over all ranks with NCCL all-to-all transposes (strong scaling of
          the FFT core, ipde_b200/distributed.py).
  --impl reference   the CPU path (oracle port of the reference's NumPy/SciPy algorithm; the reference itself
          cannot run: pybie2d/qfs/near_finder/finufft are not installable here) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

HELMHOLTZ_K = 2.0
ALG_BYTES_PER_POINT_CORE = 64.0      # scalar core, 4 sweeps of 16 B


def nice_size(n):
    """smallest 2^a 3^b 5^c >= n (FFT-friendly boundary point count)"""
    n = int(np.ceil(n))
    while True:
        m = n
        for p in (2, 3, 5):
            while m % p == 0:
                m //= p
        if m == 1:
            return n
        n += 1


def build_problem(ng, M=16):
    """interior_modified_helmholtz.py geometry with the boundary spacing matched to the grid spacing"""
    from ipde_b200.geometry import GSB, Grid, star
    from ipde_b200.embedded_boundary import EmbeddedBoundary
    from ipde_b200.ebdy_collection import EmbeddedBoundaryCollection
    from ipde_b200.embedded_function import EmbeddedFunction
    from ipde_b200.slepian import SlepianMollifier
    h = 2.4/ng
    # boundary spacing matched to the grid spacing (min speed of the a=0.1 star is 0.9: bh = h at nb = 2.36 ng);
    # nb = 2 ng (bh = 1.18 h) keeps every FFT length of the path a power of two
    nb = 2*ng if (ng & (ng - 1)) == 0 else nice_size(2*np.pi*0.9/h)
    bdy = GSB(c=star(nb, a=0.1, f=5))
    bh = bdy.dt*bdy.speed.min()
    grid = Grid([-1.2, 1.2], ng, [-1.2, 1.2], ng)
    MOL = SlepianMollifier(2*M)
    ebdy = EmbeddedBoundary(bdy, True, M, bh, heaviside=MOL.step)
    ebdyc = EmbeddedBoundaryCollection([ebdy])
    ebdyc.register_grid(grid)
    kk = np.pi/3
    sol = lambda x, y: np.exp(np.sin(kk*x))*np.sin(kk*y)
    force = lambda x, y: HELMHOLTZ_K**2*sol(x, y) - \
        kk**2*np.exp(np.sin(kk*x))*np.sin(kk*y)*(np.cos(kk*x)**2 - np.sin(kk*x) - 1.0)
    f = EmbeddedFunction(ebdyc, function=force)
    return ebdyc, f, nb


class ClockSampler(object):
    def __init__(self, index):
        self.path = '/tmp/ipde_clocks_%d_%d.csv' % (os.getpid(), index)
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.f = open(self.path, 'w')
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + q,
                                       '--format=csv,noheader,nounits', '-lms', '50'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
            # nvidia-smi needs a moment to start: wait for its first sample so that short timed regions are covered
            t0 = time.time()
            while time.time() - t0 < 5.0 and os.path.getsize(self.path) == 0:
                time.sleep(0.05)
        except Exception:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        if sm:
            out = {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(np.max(mx)), 'reasons': sorted(reasons)}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


def cpu_solve_rate(ng_sample, steps=1):
    """CPU port (oracle) of the reference path on a bounded sample of the workload."""
    from oracle.solvers import OracleScalarSolver
    from ipde_b200.qfs import Modified_Helmholtz_QFS
    ebdyc, f, nb = build_problem(ng_sample, M=16)
    e = ebdyc[0]
    pairs = [(Modified_Helmholtz_QFS(e.interface, True, True, True, HELMHOLTZ_K, tol=1e-12),
              Modified_Helmholtz_QFS(e.interface, False, True, True, HELMHOLTZ_K, tol=1e-12))]
    S = OracleScalarSolver(ebdyc, 'modified_helmholtz', HELMHOLTZ_K, 'spectral', pairs)
    t0 = time.perf_counter()
    for _ in range(steps):
        S(np.asarray(f), tol=1e-12, restart=50, maxiter=10)
    dt = (time.perf_counter() - t0)/steps
    return ng_sample*ng_sample/dt, dt, nb


def run_reference(args, rank, world):
    if rank != 0:
        return
    ng_s = args.cpu_sample_n
    steps = max(1, min(args.steps, 2))
    for _ in range(min(args.warmup, 1)):
        cpu_solve_rate(max(64, ng_s//2), 1)
    rate, dt, nb = cpu_solve_rate(ng_s, steps)
    cores = os.cpu_count()
    sample = ('same star / k=2 / M=16 problem at grid %d^2 (nb=%d), NumPy/SciPy port of the reference algorithm '
              '(direct NUDFT + direct sums), numpy/scipy threading as installed' % (ng_s, nb))
    line = {'impl': 'reference', 'metric': 'grid-point-solves/sec', 'value': rate, 'unit': 'grid-point-solves/s',
            'n_gpus': args.gpus, 'steps': steps, 'warmup': min(args.warmup, 1), 'ms_per_step': dt*1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'interior_modified_helmholtz %d^2 (CPU sample %d^2)' % (args.n, ng_s),
                       'helmholtz_k': HELMHOLTZ_K, 'M': 16},
            'cpu_baseline': {'value': rate, 'unit': 'grid-point-solves/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': rate, 'unit': 'grid-point-solves/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def ncu_traffic(kernel):
    """dram read+write bytes per launch of `kernel` from the committed ncu --set full summary (profiles/), or None"""
    path = os.path.join(ROOT, 'profiles', 'r01_ncu_%s.json' % kernel)
    try:
        recs = json.load(open(path))
    except Exception:
        return None
    tot = []
    for r in recs:
        try:
            def val(key):
                v, u = r[key].split()[:2]
                return float(v)*{'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
            tot.append(val('dram__bytes_read.sum') + val('dram__bytes_write.sum'))
        except Exception:
            pass
    return float(np.mean(tot)) if tot else None


def parse_profile(text):
    recs = []
    for line in text.strip().splitlines():
        parts = line.split()
        if len(parts) == 4:
            recs.append({'kernel': parts[0], 'launches': int(parts[1]), 'ms': float(parts[2]), 'bytes': float(parts[3])})
    return recs


def run_slab(args, world, rank, barrier):
    """periodic grid solve of ONE grid slab-decomposed over all ranks (strong scaling of the FFT core)"""
    import torch
    import torch.distributed as dist
    from ipde_b200.distributed import SlabGridSolver
    ns = args.slab_grid
    S = SlabGridSolver(ns, ns, 2.4/ns, 2.4/ns, 1, HELMHOLTZ_K, chunks=args.slab_chunks)
    gen = torch.Generator(device='cuda')
    gen.manual_seed(rank)
    fl = torch.randn((S.L.rows, ns), dtype=torch.float64, device='cuda', generator=gen)
    out = torch.empty_like(fl)
    for _ in range(3):
        S.solve(fl, out=out)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.steps):
        S.solve(fl, out=out)
    s1.record()
    barrier()
    tt = torch.tensor([s0.elapsed_time(s1)/args.steps], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return {'what': 'periodic grid solve (r2c rows, all-to-all, column FFT + symbol + inverse, all-to-all, c2r rows) '
                    'of ONE %d^2 grid slab-decomposed over all ranks; strong scaling of the FFT core' % ns,
            'grid': [ns, ns], 'n_gpus': world, 'chunks': S.K, 'ms_per_solve': float(tt[0]),
            'grid_point_solves_per_s': ns*ns/(float(tt[0])*1e-3),
            'alltoall_bytes_per_rank_per_solve': int(S.alltoall_bytes_per_rank)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ipde_b200')
    ap.add_argument('--grid', '--n', dest='n', type=int, default=2048, help='background grid points per side')
    ap.add_argument('--cpu-sample-n', type=int, default=256)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--slab-grid', type=int, default=16384,
                    help='grid size of the slab-decomposed (all ranks, one problem) grid-solve timing; 0 = skip')
    ap.add_argument('--slab-chunks', type=int, default=4,
                    help='sub-blocks of the column slab whose all-to-all overlaps the column pass (ranks > 1)')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import ctypes
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from ipde_b200._lib import lib
    from ipde_b200.device import to_dev, StageTimer, stream
    from ipde_b200.embedded_function import EmbeddedFunction
    from ipde_b200.solvers.multi_boundary.modified_helmholtz import ModifiedHelmholtzSolver

    t_setup = time.perf_counter()
    ebdyc, f, nb = build_problem(args.n)
    solver = ModifiedHelmholtzSolver(ebdyc, HELMHOLTZ_K, solver_type='spectral', grid_backend='ewald')
    t_setup = time.perf_counter() - t_setup
    gmres = dict(tol=1e-12, restart=50, maxiter=10)
    # every rank its own forcing (weak scaling): same geometry, scaled right-hand side
    f_host = np.asarray(f)*(1.0 + 0.1*rank)
    f_dev = to_dev(f_host)
    npts = args.n*args.n
    warmup = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        solver.solve_device(f_dev, **gmres)
    barrier()

    # ---- device-resident throughput (value), with per-stage events and clocks
    sampler = ClockSampler(local_rank)
    timer = StageTimer()
    solver.timer = timer
    launches0 = lib().ipde_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        res = solver.solve_device(f_dev, **gmres)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = lib().ipde_launch_count() - launches0
    stages = timer.results()
    solver.timer = None

    # ---- end to end through the public API: host EmbeddedFunction in, host EmbeddedFunction out
    barrier()
    f_ef = EmbeddedFunction(ebdyc, linear_data=f_host)
    solver(f_ef, **gmres)
    barrier()
    ue = solver(f_ef, **gmres)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ue = solver(f_ef, out=ue, **gmres)          # host in, host out; the result array is reused across steps
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()

    # ---- per-kernel CUDA-event times of one more step (library profiler: an event after every launch)
    prof_steps = 3
    lib().ipde_profile_start(stream())
    for _ in range(prof_steps):
        solver.solve_device(f_dev, **gmres)
    buf = ctypes.create_string_buffer(1 << 16)
    lib().ipde_profile_stop(buf, len(buf))
    prof = parse_profile(buf.value.decode())
    gmres_its = solver.iteration_counts

    times = torch.tensor([ms_total, e2e_s*1e3], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    # ---- one problem over all ranks: slab-decomposed periodic grid solve (NCCL all-to-all transposes)
    slab = None
    if args.slab_grid > 0 and args.slab_grid % world == 0:
        del solver, res
        torch.cuda.empty_cache()
        try:
            slab = run_slab(args, world, rank, barrier)
        except Exception as exc:          # the headline numbers above are already reduced over the ranks
            slab = {'error': '%s: %s' % (type(exc).__name__, exc)}
    ms_total, e2e_ms = float(times[0]), float(times[1])
    ms_per_step = ms_total/args.steps
    value = world*npts*args.steps/(ms_total*1e-3)
    e2e_value = world*npts*args.steps/(e2e_ms*1e-3)

    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        per_step = {k: v/args.steps for k, v in stages.items()}
        prof_ms = sum(r['ms'] for r in prof)
        tagged = [r for r in prof if r['bytes'] > 0]
        dom = max(tagged, key=lambda r: r['ms'])
        achieved = dom['bytes']/(dom['ms']*1e-3)/1e9
        roofline = {'bound': 'hbm', 'kernel': dom['kernel'], 'achieved': achieved, 'peak': peak,
                    'peak_source': peak_kind, 'unit': 'GB/s', 'frac': achieved/peak, 'traffic': ncu_traffic(dom['kernel']),
                    'algorithmic_bytes_per_launch': dom['bytes']/dom['launches'],
                    'ms_per_launch': dom['ms']/dom['launches'], 'launches_per_step': dom['launches']/prof_steps,
                    'share_of_step': dom['ms']/prof_ms,
                    'how': 'CUDA event after every launch of libipde_b200 on the solve stream (ipde_profile_start/'
                           'stop), %d profiled steps after the timed region' % prof_steps,
                    'kernels_ms_per_step': {r['kernel']: r['ms']/prof_steps for r in prof[:12]},
                    'kernels_gbs': {r['kernel']: r['bytes']/(r['ms']*1e-3)/1e9 for r in tagged[:10]},
                    'stage_ms_per_step': per_step,
                    'whole_step_algorithmic': {
                        'bytes_per_point': ALG_BYTES_PER_POINT_CORE,
                        'note': 'SURVEY 8(d) scalar core only; evaluator/NUFFT passes come on top (DESIGN.md)',
                        'frac_of_peak': ALG_BYTES_PER_POINT_CORE*npts/(ms_per_step*1e-3)/1e9/peak}}
        line = {'metric': 'grid-point-solves/sec', 'value': value, 'unit': 'grid-point-solves/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': warmup, 'ms_per_step': ms_per_step,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
                'data': 'synthetic',
                'config': {'workload': 'interior_modified_helmholtz (BASELINE configs[1]) %d^2 grid, 1 star boundary '
                           'nb=%d, M=16, k=2, spectral interface interpolation, annular GMRES tol 1e-12, '
                           'spectral-Ewald evaluation of the layer sums' % (args.n, nb),
                           'grid': [args.n, args.n], 'n_boundary': nb, 'dof': int(ebdyc.dof),
                           'gmres_iterations': gmres_its, 'setup_s': t_setup,
                           'l2': 'inputs larger than L2: every step streams ~%.1f GB of grid-sized arrays '
                                 '(126 MB L2); no flush between steps' % (npts*8*40/1e9)
                                 if npts*8*40 > 4*126e6 else 'working set comparable to L2 (126 MB) at this size'},
                'clocks': clocks, 'gpu_launches': int(launches),
                'e2e': {'value': e2e_value, 'unit': 'grid-point-solves/s',
                        'h2d_bytes_per_step': int(f_host.nbytes), 'd2h_bytes_per_step': int(np.asarray(ue).nbytes),
                        'ms_per_step': e2e_ms/args.steps,
                        'api': 'ModifiedHelmholtzSolver.__call__(EmbeddedFunction, out=EmbeddedFunction) -> EmbeddedFunction: '
                               'host arrays in and out through pinned staging buffers, result array reused across steps'},
                'roofline': roofline}
        if slab is not None:
            line['slab_grid_solve'] = slab
        if not args.no_cpu_baseline and world == 1:
            rate, dt, nbs = cpu_solve_rate(args.cpu_sample_n, 1)
            line['cpu_baseline'] = {'value': rate, 'unit': 'grid-point-solves/s', 'cores': os.cpu_count(),
                                    'kind': 'port',
                                    'sample': 'one solve of the same problem at grid %d^2 (nb=%d) by the NumPy/SciPy '
                                              'oracle port (direct NUDFT + all-pairs sums, as the reference\'s default '
                                              'backend), %.1f s' % (args.cpu_sample_n, nbs, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
