"""Slab decomposition of the grid solve: index logic and exchanges on CPU (gloo, world_size 2 and 3); the device
kernels on one GPU (world 1) in the gpu-marked test; multi-GPU runs go through scripts/run_slab_multi.py."""
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from ipde_b200.distributed import SlabLayout, slab_grid_solve_numpy, slab_p2p_solve_virtual, p2p_columns_per_rank


def _reference(f, xh, yh, k):
    nx, ny = f.shape
    kx = np.fft.fftfreq(nx, xh/(2*np.pi))[:, None]
    ky = np.fft.fftfreq(ny, yh/(2*np.pi))[None, :]
    return np.fft.ifft2(np.fft.fft2(f)/(k*k + kx*kx + ky*ky)).real


def _worker(rank, world, port, nx, ny, q, chunks=1):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        f = rng.standard_normal((nx, ny))
        xh, yh, k = 2.4/nx, 2.4/ny, 2.0
        L = SlabLayout(nx, ny, world, rank)
        def symbol(i, j):
            kx = np.fft.fftfreq(nx, xh/(2*np.pi))[i]
            ky = 2*np.pi*j/(ny*yh)
            return 1.0/(k*k + kx*kx + ky*ky)
        u = slab_grid_solve_numpy(L, f[L.row0:L.row0 + L.rows], symbol, chunks=chunks)
        ref = _reference(f, xh, yh, k)[L.row0:L.row0 + L.rows]
        q.put((rank, float(np.abs(u - ref).max()/np.abs(ref).max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,nx,ny,chunks', [(2, 32, 48, 1), (3, 36, 40, 1), (2, 32, 48, 3), (3, 36, 64, 2)])
def test_slab_solve_gloo(world, nx, ny, chunks):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + world + 10*chunks
    procs = [ctx.Process(target=_worker, args=(r, world, port, nx, ny, q, chunks)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert len(res) == world
    for rank, err in res:
        assert err < 1e-13, (rank, err)


def test_slab_pack_unpack_roundtrip():
    rng = np.random.default_rng(1)
    for world, nx, ny in [(1, 8, 10), (2, 8, 10), (4, 16, 22), (8, 16, 8)]:
        L = SlabLayout(nx, ny, world, 0)
        s = rng.standard_normal((L.rows, L.P)) + 1j*rng.standard_normal((L.rows, L.P))
        assert np.array_equal(L.unpack(L.pack(s)), s)
        assert L.Pc*world >= L.P
        # chunked: the sub-blocks tile every rank's block, widths rounded up
        for K in (2, 3):
            Lc = SlabLayout(nx, ny, world, 0)
            Pk = Lc.set_chunks(K, round_up=lambda n: n + (n % 2))
            assert Lc.Pc == Pk*K and Lc.Pc*world >= Lc.P
            back = np.zeros_like(s)
            for c in range(K):
                Lc.unpack_chunk(Lc.pack_chunk(s, c*Pk, Pk), c*Pk, Pk, back)
            assert np.array_equal(back, s)


@pytest.mark.parametrize('world,nx,ny', [(2, 16, 12), (4, 32, 50), (8, 64, 30), (3, 24, 16), (8, 8, 64)])
def test_slab_p2p_addresses_numpy(world, nx, ny):
    """host statement of the fused transposes (the address functions the CUDA stores implement, csrc/fft.cuh) for all
    ranks in one process: every spectrum entry lands exactly once, the result is the unsharded solve"""
    rng = np.random.default_rng(world*nx + ny)
    f = rng.standard_normal((nx, ny))
    xh, yh, k = 2.4/nx, 3.1/ny, 2.0
    def symbol(i, j):
        kx = np.fft.fftfreq(nx, xh/(2*np.pi))[i]
        ky = 2*np.pi*j/(ny*yh)
        return 1.0/(k*k + kx*kx + ky*ky)
    round8 = lambda n: -(-n//8)*8          # any rounding of the column width must leave the result unchanged
    for nxt in (lambda n: n, round8):
        assert world*p2p_columns_per_rank(ny//2 + 1, world, nxt) >= ny//2 + 1
        u = slab_p2p_solve_virtual(nx, ny, world, f, symbol, nxt)
        ref = _reference(f, xh, yh, k)
        assert np.abs(u - ref).max()/np.abs(ref).max() < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(1024, 1024), (2048, 8192), (1024, 16384), (1024, 32768)])
def test_slab_p2p_solver_single_gpu(shape):
    """SlabGridSolverP2P on one rank: the transposing stores of the row and column kernels land in the rank's own IPC
    buffers (register r2c kernel, long-row untangle kernel, short four-step column pass)"""
    from ipde_b200.distributed import SlabGridSolverP2P
    import ipde_b200.device as D
    nx, ny = shape
    rng = np.random.default_rng(5)
    f = rng.standard_normal((nx, ny))
    xh, yh = 2.4/nx, 3.0/ny
    S = SlabGridSolverP2P(nx, ny, xh, yh, 1, 2.0)
    u = S.solve(D.to_dev(f)).cpu().numpy()
    S.close()
    ref = _reference(f, xh, yh, 2.0)
    assert np.abs(u - ref).max()/np.abs(ref).max() < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize('R,shape', [(2, (1024, 1024)), (4, (2048, 4096)), (8, (1024, 16384)), (8, (2048, 32768))])
def test_slab_p2p_virtual_ranks(R, shape):
    """the index arithmetic of the fused transposes for R ranks, all of them played by one GPU: R column slabs and R row
    slabs in one process, the C entry points called rank after rank in the order of SlabGridSolverP2P.solve"""
    import ctypes
    from ipde_b200._lib import lib
    import ipde_b200.device as D
    nx, ny = shape
    if nx % R:
        pytest.skip('nx not divisible')
    rng = np.random.default_rng(R)
    f = rng.standard_normal((nx, ny))
    xh, yh = 2.4/nx, 3.0/ny
    L_ = lib()
    layouts = [SlabLayout(nx, ny, R, r) for r in range(R)]
    Pc = L_.ipde_fft_next_size(max(2, layouts[0].Pc - 1)) + 1
    rows, P = layouts[0].rows, layouts[0].P
    if nx < 1024 or nx & (nx - 1):
        pytest.skip('column length without a four-step split')
    row_plan, col_plan = ctypes.c_void_p(), ctypes.c_void_p()
    L_.ipde_fft_plan_create(ny//2, ctypes.byref(row_plan))
    L_.ipde_grid_plan_create(nx, 2*(Pc - 1), xh, yh, ctypes.byref(col_plan))
    cols = [torch.zeros((nx, Pc), dtype=torch.complex128, device='cuda') for _ in range(R)]
    back = [torch.zeros((rows, P), dtype=torch.complex128, device='cuda') for _ in range(R)]
    peer_cols = (ctypes.c_void_p*R)(*[D.ptr(c) for c in cols])
    peer_rows = (ctypes.c_void_p*R)(*[D.ptr(b) for b in back])
    scratch, work = D.empty((rows, P), torch.complex128), D.empty((nx, Pc), torch.complex128)
    st = D.stream()
    for r in range(R):
        fl = D.to_dev(f[r*rows:(r + 1)*rows])
        L_.ipde_slab_r2c_rows_scatter(row_plan, D.ptr(fl), D.ptr(scratch), rows, ny, P, R, Pc, r*rows, peer_cols, st)
    spec = np.fft.rfft(f, axis=1)
    for r in range(R):
        a, b = r*Pc, min(P, (r + 1)*Pc)
        assert np.abs(cols[r].cpu().numpy()[:, :b - a] - spec[:, a:b]).max() < 1e-12*np.abs(spec).max()
    for r in range(R):
        L_.ipde_slab_cols_symbol_scatter(col_plan, D.ptr(cols[r]), D.ptr(work), 1, 2.0, ny, yh, r*Pc, R, rows, P, peer_rows, st)
    ref = _reference(f, xh, yh, 2.0)
    for r in range(R):
        out = D.empty((rows, ny))
        L_.ipde_fft_c2r_rows(row_plan, D.ptr(back[r]), D.ptr(out), rows, P, ny, 2.0/(nx*ny), st)
        assert np.abs(out.cpu().numpy() - ref[r*rows:(r + 1)*rows]).max()/np.abs(ref).max() < 1e-13
    L_.ipde_fft_plan_destroy(row_plan)
    L_.ipde_grid_plan_destroy(col_plan)


@pytest.mark.gpu
@pytest.mark.parametrize('n', [256, 4096])
def test_slab_grid_solver_single_gpu(n):
    """device kernels of the slab path (world 1: pack/unpack, column-slab symbol pass) against NumPy"""
    from ipde_b200.distributed import SlabGridSolver
    import ipde_b200.device as D
    rng = np.random.default_rng(2)
    f = rng.standard_normal((n, n))
    xh = 2.4/n
    S = SlabGridSolver(n, n, xh, xh, 1, 2.0)
    u = S.solve(D.to_dev(f)).cpu().numpy()
    ref = _reference(f, xh, xh, 2.0)
    assert np.abs(u - ref).max()/np.abs(ref).max() < 1e-13
    # CUDA pack/unpack against the NumPy statement for a 4-rank layout
    from ipde_b200._lib import lib
    L = SlabLayout(64, 50, 4, 1)
    s = rng.standard_normal((L.rows, L.P)) + 1j*rng.standard_normal((L.rows, L.P))
    ds = D.to_dev(s)
    packed = D.empty((L.R, L.rows, L.Pc), torch.complex128)
    lib().ipde_slab_pack(D.ptr(ds), L.rows, L.P, L.R, L.Pc, D.ptr(packed), D.stream())
    assert np.array_equal(packed.cpu().numpy(), L.pack(s))
    back = D.empty((L.rows, L.P), torch.complex128)
    lib().ipde_slab_unpack(D.ptr(packed), L.rows, L.P, L.R, L.Pc, D.ptr(back), D.stream())
    assert np.array_equal(back.cpu().numpy(), s)
    # chunked variants
    Lc = SlabLayout(64, 50, 4, 1)
    Pk = Lc.set_chunks(2, lib().ipde_fft_next_size)
    back = torch.zeros((Lc.rows, Lc.P), dtype=torch.complex128, device='cuda')
    for c in range(2):
        pc = D.empty((Lc.R, Lc.rows, Pk), torch.complex128)
        lib().ipde_slab_pack_chunk(D.ptr(ds), Lc.rows, Lc.P, Lc.R, Lc.Pc, c*Pk, Pk, D.ptr(pc), D.stream())
        assert np.array_equal(pc.cpu().numpy(), Lc.pack_chunk(s, c*Pk, Pk))
        lib().ipde_slab_unpack_chunk(D.ptr(pc), Lc.rows, Lc.P, Lc.R, Lc.Pc, c*Pk, Pk, D.ptr(back), D.stream())
    assert np.array_equal(back.cpu().numpy(), s)


def _assign_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from types import SimpleNamespace as NS
        from ipde_b200.solvers.multi_boundary.scalar import ScalarSolver
        from ipde_b200.solvers.multi_boundary.vector import VectorSolver
        ebdyc = [NS(bdy=NS(N=n), M=m) for n, m in ((800, 16), (200, 16), (200, 16), (120, 12))]
        out = []
        for cls in (ScalarSolver, VectorSolver):
            s = NS(ebdyc=ebdyc)
            cls._assign_boundaries(s, True)
            out.append((s.world, s.rank, tuple(s.owner)))
            s2 = NS(ebdyc=ebdyc)
            cls._assign_boundaries(s2, False)
            out.append((s2.world, s2.rank, tuple(s2.owner)))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_boundary_sharding_assignment_gloo():
    """greedy cost-weighted assignment of boundaries to ranks: identical on every rank, the big interior
    boundary alone on one rank; without sharding everything stays on the calling rank"""
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + 77
    procs = [ctx.Process(target=_assign_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for rank in range(world):
        for k, (w, r, owner) in enumerate(res[rank]):
            if k % 2 == 0:
                assert (w, r) == (world, rank) and owner == (0, 1, 1, 1)
            else:
                assert (w, r) == (1, 0) and owner == (0, 0, 0, 0)
    assert res[0][0][2] == res[1][0][2]
