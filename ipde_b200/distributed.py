"""
Slab-decomposed periodic grid solve across the GPUs of one box.

The reference has no multi-process code; what is distributed here is its `_grid_solve`
(ipde/solvers/multi_boundary/modified_helmholtz.py:40-46, poisson.py:30-38): rank p owns rows
[p nx/R, (p+1) nx/R) of every (nx, ny) grid field.

    local real rows -> half-complex rows        k_r2c_rows
    pack column blocks                          k_slab_pack            (rows, P) -> (R, rows, Pc)
    all-to-all                                  NCCL over NVLink (torch.distributed.all_to_all_single)
    column FFT + symbol + inverse column FFT    k_fft_cols / k_cols_mult on the (nx, Pc) column slab
    all-to-all back                             NCCL
    unpack, half-complex rows -> real rows      k_slab_unpack, k_c2r_rows

Two all-to-alls per scalar solve, each moving 16 B * (nx/R) * Pc to every peer.  The received blocks are already the
column slab in row order (rank q' sends rows q' nx/R ...), so only the sending side of the forward transpose and the
receiving side of the backward one need a pack/unpack pass.

`SlabLayout` holds the index arithmetic with NumPy stand-ins for the local transforms, so the partition/exchange
logic is tested on CPU with the gloo backend (tests/test_distributed.py); `SlabGridSolver` is the device path.
"""
import numpy as np
import torch
import torch.distributed as dist


class SlabLayout(object):
    def __init__(self, nx, ny, world, rank):
        if nx % world:
            raise Exception('slab decomposition needs nx divisible by the number of ranks')
        self.nx, self.ny, self.R, self.rank = nx, ny, world, rank
        self.rows = nx//world
        self.row0 = rank*self.rows
        self.P = ny//2 + 1
        self.Pc = -(-self.P//world)
        if self.Pc < 3:
            self.Pc = 3
        self.col0 = rank*self.Pc

    # NumPy statements of the pack / unpack passes (the CUDA kernels are checked against these)
    def pack(self, spec_rows):
        out = np.zeros((self.R, self.rows, self.Pc), dtype=complex)
        for q in range(self.R):
            a, b = q*self.Pc, min(self.P, (q + 1)*self.Pc)
            if b > a:
                out[q, :, :b - a] = spec_rows[:, a:b]
        return out

    def unpack(self, blocks):
        out = np.empty((self.rows, self.P), dtype=complex)
        for q in range(self.R):
            a, b = q*self.Pc, min(self.P, (q + 1)*self.Pc)
            if b > a:
                out[:, a:b] = blocks[q, :, :b - a]
        return out

    # sub-block [koff, koff + Pk) of every rank's column block (chunked, overlapped exchange)
    def set_chunks(self, K, round_up=lambda n: n):
        """make Pc a multiple of K chunks of width Pk, with Pk - 1 rounded up by `round_up` (a supported FFT length)"""
        Pk = round_up(max(2, -(-self.Pc//K) - 1)) + 1
        self.Pc = Pk*K
        self.col0 = self.rank*self.Pc
        return Pk

    def pack_chunk(self, spec_rows, koff, Pk):
        out = np.zeros((self.R, self.rows, Pk), dtype=complex)
        for q in range(self.R):
            a = q*self.Pc + koff
            b = min(self.P, a + Pk)
            if b > a:
                out[q, :, :b - a] = spec_rows[:, a:b]
        return out

    def unpack_chunk(self, blocks, koff, Pk, out):
        for q in range(self.R):
            a = q*self.Pc + koff
            b = min(self.P, a + Pk)
            if b > a:
                out[:, a:b] = blocks[q, :, :b - a]
        return out


def exchange(send, recv, group=None):
    """all-to-all of equal contiguous blocks along dim 0; NCCL collective on the device, point-to-point on gloo"""
    if dist.get_backend(group) == 'nccl':
        dist.all_to_all_single(recv, send, group=group)
        return
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    reqs = []
    for q in range(world):
        if q == rank:
            recv[q].copy_(send[q])
        else:
            reqs.append(dist.isend(send[q].contiguous(), q, group=group))
            reqs.append(dist.irecv(recv[q], q, group=group))
    for r in reqs:
        r.wait()


def slab_grid_solve_numpy(layout, f_local, symbol, group=None, chunks=1):
    """CPU statement of the distributed solve (NumPy local transforms + the same exchanges): returns the local rows of
    Re ifft2(symbol * fft2(f)).  `symbol(kx_index_array, ky_index_array)` gives the real multiplier.  chunks > 1
    follows the sub-block schedule of SlabGridSolver (one exchange per sub-block each way)."""
    L = layout
    spec = np.fft.rfft(f_local, axis=1)                                   # (rows, P)
    if chunks > 1:
        Pk = L.set_chunks(chunks)
        back_spec = np.zeros((L.rows, L.P), dtype=complex)
        for c in range(chunks):
            send = torch.from_numpy(L.pack_chunk(spec, c*Pk, Pk))
            recv = torch.empty_like(send)
            exchange(send, recv, group)
            cols = recv.numpy().reshape(L.nx, Pk)
            j = L.col0 + c*Pk + np.arange(Pk)
            valid = j < L.P
            m = np.zeros((L.nx, Pk))
            m[:, valid] = symbol(np.arange(L.nx)[:, None], j[valid][None, :])
            back = np.fft.ifft(np.fft.fft(cols, axis=0)*m, axis=0)
            send2 = torch.from_numpy(np.ascontiguousarray(back.reshape(L.R, L.rows, Pk)))
            recv2 = torch.empty_like(send2)
            exchange(send2, recv2, group)
            L.unpack_chunk(recv2.numpy(), c*Pk, Pk, back_spec)
        return np.fft.irfft(back_spec, n=L.ny, axis=1)
    send = torch.from_numpy(L.pack(spec))
    recv = torch.empty_like(send)
    exchange(send, recv, group)
    cols = recv.numpy().reshape(L.nx, L.Pc)                               # column slab, rows in global order
    ch = np.fft.fft(cols, axis=0)
    j = L.col0 + np.arange(L.Pc)
    valid = j < L.P
    m = np.zeros((L.nx, L.Pc))
    m[:, valid] = symbol(np.arange(L.nx)[:, None], j[valid][None, :])
    back = np.fft.ifft(ch*m, axis=0)
    send2 = torch.from_numpy(np.ascontiguousarray(back.reshape(L.R, L.rows, L.Pc)))
    recv2 = torch.empty_like(send2)
    exchange(send2, recv2, group)
    return np.fft.irfft(L.unpack(recv2.numpy()), n=L.ny, axis=1)


class SlabGridSolver(object):
    """
    Device path.  kind: 0 Poisson (zero mode -> 0; the caller demeans), 1 modified Helmholtz.
    solve(f_local (rows, ny) float64 cuda) -> u_local (rows, ny).

    chunks > 1 (ranks > 1 only): every rank's column block is cut into `chunks` sub-blocks; the all-to-all of
    sub-block k+1 (NCCL's own stream, async_op) runs while the column pass of sub-block k does, and the way back
    likewise, so only the first pack and the last unpack are exposed next to the row transforms.
    """
    def __init__(self, nx, ny, xh, yh, kind, k, group=None, chunks=1):
        import ctypes
        from ._lib import lib
        from .device import require_cuda, empty, C128
        require_cuda()
        self.group = group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.L = SlabLayout(nx, ny, world, rank)
        Ls = self.L
        K = max(1, int(chunks)) if world > 1 else 1
        while K > 1 and -(-Ls.Pc//K) < 3:
            K -= 1
        self.K = K
        if world > 1:
            # the column plan is a grid plan of width 2 (Pk - 1): round the chunk width up to a supported length
            # (a few zero columns), and make Pc a multiple of the number of chunks
            Ls.set_chunks(K, lib().ipde_fft_next_size)
        self.Pk = Ls.Pc//K
        self.nx, self.ny, self.xh, self.yh, self.kind, self.k = nx, ny, float(xh), float(yh), int(kind), float(k)
        L_ = lib()
        self.row_plan = ctypes.c_void_p()
        L_.ipde_fft_plan_create(ny//2, ctypes.byref(self.row_plan))
        self.col_plan = ctypes.c_void_p()
        L_.ipde_grid_plan_create(nx, 2*(self.Pk - 1), float(xh), float(yh), ctypes.byref(self.col_plan))
        self.spec = empty((Ls.rows, Ls.P), C128)
        self.send = [empty((Ls.R, Ls.rows, self.Pk), C128) for _ in range(K)]
        self.recv = [empty((Ls.R, Ls.rows, self.Pk), C128) for _ in range(K)]
        self.work = [empty((Ls.R, Ls.rows, self.Pk), C128) for _ in range(K)]
        self.alltoall_bytes_per_rank = 2*16*Ls.rows*Ls.Pc*(Ls.R - 1)

    def __del__(self):
        try:
            from ._lib import lib
            lib().ipde_fft_plan_destroy(self.row_plan)
            lib().ipde_grid_plan_destroy(self.col_plan)
        except Exception:
            pass

    def _exchange(self, send, recv):
        """all-to-all; on NCCL asynchronously (returns the work handle, None otherwise)"""
        if dist.get_backend(self.group) == 'nccl':
            return dist.all_to_all_single(recv, send, group=self.group, async_op=True)
        exchange(send, recv, self.group)
        return None

    def solve(self, f_local, out=None):
        from ._lib import lib
        from .device import ptr, stream, empty
        L_, Ls, st = lib(), self.L, stream()
        if out is None:
            out = empty((Ls.rows, self.ny))
        L_.ipde_fft_r2c_rows(self.row_plan, ptr(f_local), ptr(self.spec), Ls.rows, self.ny, Ls.P, st)
        if Ls.R == 1:
            # one rank: the row spectrum already is the column slab (Pc == P), no pack
            L_.ipde_grid_cols_symbol_slab(self.col_plan, ptr(self.spec), ptr(self.work[0]), 0, 0, self.kind, self.k,
                                          self.ny, self.yh, 0, st)
            back = self.work[0]
        else:
            K, Pk = self.K, self.Pk
            fwd = []
            for c in range(K):
                L_.ipde_slab_pack_chunk(ptr(self.spec), Ls.rows, Ls.P, Ls.R, Ls.Pc, c*Pk, Pk, ptr(self.send[c]), st)
                fwd.append(self._exchange(self.send[c], self.recv[c]))
            bwd = []
            for c in range(K):
                if fwd[c] is not None:
                    fwd[c].wait()
                # column sub-slab (nx, Pk): forward columns, symbol, inverse columns; the received blocks are already
                # in global row order
                L_.ipde_grid_cols_symbol_slab(self.col_plan, ptr(self.recv[c]), ptr(self.work[c]), 0, 0, self.kind,
                                              self.k, self.ny, self.yh, Ls.col0 + c*Pk, st)
                bwd.append(self._exchange(self.work[c], self.send[c]))
            for c in range(K):
                if bwd[c] is not None:
                    bwd[c].wait()
                L_.ipde_slab_unpack_chunk(ptr(self.send[c]), Ls.rows, Ls.P, Ls.R, Ls.Pc, c*Pk, Pk, ptr(self.spec), st)
            back = self.spec
        L_.ipde_fft_c2r_rows(self.row_plan, ptr(back), ptr(out), Ls.rows, Ls.P, self.ny,
                             2.0/(float(self.nx)*float(self.ny)), st)
        return out


def p2p_columns_per_rank(P, world, next_size=lambda n: n):
    """columns per rank of SlabGridSolverP2P: ceil(P / world), with Pc - 1 rounded up to a supported FFT length (the
    column plan is a grid plan of width 2 (Pc - 1)); world * Pc >= P always"""
    Pc = max(3, -(-P//world))
    return next_size(max(2, Pc - 1)) + 1


def p2p_forward_address(row0, l, k, Pc):
    """entry (local row l, column k) of the row spectra of the rank owning rows row0... -> (destination rank, offset in
    its column slab (nx, Pc)); csrc/fft.cuh:slab_fwd_addr"""
    q = k//Pc
    return q, (row0 + l)*Pc + (k - q*Pc)


def p2p_backward_address(i, c, rows, P, col0):
    """entry (global row i, local column c) of the rank owning columns col0... -> (destination rank, offset in its row
    slab (rows, P)); only columns col0 + c < P are stored; csrc/fft.cuh:slab_bwd_addr"""
    q = i//rows
    return q, (i - q*rows)*P + col0 + c


def slab_p2p_solve_virtual(nx, ny, world, f, symbol, next_size=lambda n: n):
    """NumPy statement of SlabGridSolverP2P with all ranks played by one process: the two transposes are the scattered
    stores of the address functions above, nothing else moves data.  Returns the full (nx, ny) result."""
    P = ny//2 + 1
    rows = nx//world
    Pc = p2p_columns_per_rank(P, world, next_size)
    cols = [np.zeros(nx*Pc, dtype=complex) for _ in range(world)]
    back = [np.zeros(rows*P, dtype=complex) for _ in range(world)]
    kk = np.arange(P)
    for r in range(world):
        spec = np.fft.rfft(f[r*rows:(r + 1)*rows], axis=1)
        for l in range(rows):
            q, off = p2p_forward_address(r*rows, l, kk, Pc)
            for dest in range(world):
                m = q == dest
                cols[dest][off[m]] = spec[l, m]
    ii = np.arange(nx)
    for r in range(world):
        slab = cols[r].reshape(nx, Pc)
        j = r*Pc + np.arange(Pc)
        valid = j < P
        mult = np.zeros((nx, Pc))
        mult[:, valid] = symbol(ii[:, None], j[valid][None, :])
        res = np.fft.ifft(np.fft.fft(slab, axis=0)*mult, axis=0)
        for c in np.flatnonzero(valid):
            q, off = p2p_backward_address(ii, c, rows, P, r*Pc)
            for dest in range(world):
                m = q == dest
                back[dest][off[m]] = res[m, c]
    return np.concatenate([np.fft.irfft(b.reshape(rows, P), n=ny, axis=1) for b in back], axis=0)


class SlabGridSolverP2P(object):
    """
    The slab solve with both transposes FUSED INTO THE TRANSFORM KERNELS: no pack, no all-to-all, no unpack.

        r2c rows, every spectrum entry stored into the column slab of the rank that owns its column
                                                        k_r2c_rows_reg / k_r2c_untangle_split, NVLink stores
        tiny all-reduce (every rank's stores have landed)
        column FFT + symbol + inverse column FFT, the last pass storing every entry into the row slab of the rank
        that owns its row                               k_fft_cols_short, k_cols_mult_short, NVLink stores
        tiny all-reduce
        c2r rows                                        k_c2r_rows_*

    The buffers that peers write into are allocated by the library (cudaMalloc), exported as CUDA IPC handles and mapped
    by every other rank of the box (`ipde_ipc_*`); the handles travel once, at construction, through
    `all_gather_object`.  The data path has no collective: the exchange is carried by the stores of the kernels that
    produce the data and overlaps their arithmetic.  One rank: the same kernels store into the rank's own buffers.
    Up to 8 ranks (one box); nx a power of two >= 1024 and ny/2 a register-kernel length (256 ... 8192, 16384).
    """
    MAX_RANKS = 8

    def __init__(self, nx, ny, xh, yh, kind, k, group=None):
        import ctypes
        from ._lib import lib
        from .device import require_cuda, empty, C128
        require_cuda()
        self.group = group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world > self.MAX_RANKS:
            raise Exception('SlabGridSolverP2P: at most %d ranks' % self.MAX_RANKS)
        self.L = Ls = SlabLayout(nx, ny, world, rank)
        L_ = lib()
        # the column plan is a grid plan of width 2 (Pc - 1): round Pc - 1 up to a supported length (zero columns)
        Ls.Pc = p2p_columns_per_rank(Ls.P, world, L_.ipde_fft_next_size)
        Ls.col0 = rank*Ls.Pc
        self.nx, self.ny, self.xh, self.yh, self.kind, self.k = nx, ny, float(xh), float(yh), int(kind), float(k)
        self.row_plan = ctypes.c_void_p()
        L_.ipde_fft_plan_create(ny//2, ctypes.byref(self.row_plan))
        self.col_plan = ctypes.c_void_p()
        L_.ipde_grid_plan_create(nx, 2*(Ls.Pc - 1), float(xh), float(yh), ctypes.byref(self.col_plan))
        self.scratch = empty((Ls.rows, Ls.P), C128)
        self.work = empty((nx, Ls.Pc), C128)
        # shared buffers: my column slab (nx, Pc) and my row slab of the way back (rows, P)
        self._own, handles = [], []
        for nbytes in (16*nx*Ls.Pc, 16*Ls.rows*Ls.P):
            dp, h = ctypes.c_void_p(), ctypes.create_string_buffer(64)
            L_.ipde_ipc_alloc(nbytes, ctypes.byref(dp), h)
            self._own.append(dp)
            handles.append(h.raw)
        self._mapped = []
        tables = [[None]*world, [None]*world]
        if world > 1:
            gathered = [None]*world
            dist.all_gather_object(gathered, handles, group=group)
        failure = None
        try:
            for q in range(world):
                for which in range(2):
                    if q == rank:
                        tables[which][q] = self._own[which].value
                    else:
                        dp = ctypes.c_void_p()
                        L_.ipde_ipc_open(gathered[q][which], ctypes.byref(dp))
                        self._mapped.append(dp)
                        tables[which][q] = dp.value
        except Exception as exc:
            failure = exc
        if world > 1:
            # every rank takes the same decision: a rank that cannot map a peer must not leave the others waiting
            ok = torch.tensor([0 if failure is not None else 1], device='cuda')
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok) == 0:
                self.close()
                raise Exception('SlabGridSolverP2P: peer mapping failed on some rank (%s)' % (failure,))
        elif failure is not None:
            raise failure
        self.peer_cols = (ctypes.c_void_p*world)(*tables[0])
        self.peer_rows = (ctypes.c_void_p*world)(*tables[1])
        self._flag = torch.zeros(1, device='cuda')
        self.p2p_bytes_per_rank = 2*16*Ls.rows*Ls.P*(world - 1)//world
        self._fence()

    def close(self):
        from ._lib import lib
        L_ = lib()
        if getattr(self, '_own', None) is None:
            return
        torch.cuda.synchronize()
        if self.L.R > 1:
            dist.barrier(group=self.group)        # nobody unmaps or frees while a peer may still store
        for dp in self._mapped:
            L_.ipde_ipc_close(dp)
        if self.L.R > 1:
            dist.barrier(group=self.group)
        for dp in self._own:
            L_.ipde_ipc_free(dp)
        self._own, self._mapped = None, []
        L_.ipde_fft_plan_destroy(self.row_plan)
        L_.ipde_grid_plan_destroy(self.col_plan)

    def __del__(self):
        # buffers that peers may still have mapped are released by close() (collective); never from the collector
        pass

    def _fence(self):
        """stream-ordered: returns (on the stream) once the kernels enqueued so far on EVERY rank have completed"""
        if self.L.R > 1:
            dist.all_reduce(self._flag, group=self.group)

    def solve(self, f_local, out=None):
        from ._lib import lib
        from .device import ptr, stream, empty
        L_, Ls, st = lib(), self.L, stream()
        if out is None:
            out = empty((Ls.rows, self.ny))
        L_.ipde_slab_r2c_rows_scatter(self.row_plan, ptr(f_local), ptr(self.scratch), Ls.rows, self.ny, Ls.P, Ls.R, Ls.Pc,
                                      Ls.row0, self.peer_cols, st)
        self._fence()
        L_.ipde_slab_cols_symbol_scatter(self.col_plan, self._own[0], ptr(self.work), self.kind, self.k, self.ny, self.yh,
                                         Ls.col0, Ls.R, Ls.rows, Ls.P, self.peer_rows, st)
        self._fence()
        L_.ipde_fft_c2r_rows(self.row_plan, self._own[1], ptr(out), Ls.rows, Ls.P, self.ny,
                             2.0/(float(self.nx)*float(self.ny)), st)
        return out
