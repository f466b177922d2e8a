"""
ScalarSolver: the inhomogeneous scalar solve over an EmbeddedBoundaryCollection, on the device.

API mirror of ipde/solvers/multi_boundary/scalar.py:7-120.  `__call__(f, **kwargs)` takes an
EmbeddedFunction (host), returns an EmbeddedFunction (host); kwargs (tol, restart, maxiter,
verbose) go to the annular GMRES.  Between the one host->device copy of `f` and the one
device->host copy of the result everything runs in libipde_b200.so:

  scatter + cutoff -> (demean) -> 2-D FFT -> multiplier -> inverse FFT            scalar.py:76-78
  interface values of u, u_x, u_y: type-2 NUFFT of [uh, ikx uh, iky uh]            scalar.py:80-88
        ('fourth': cubic interpolation of uc and its 4th-order differences          scalar.py:89-95)
  per boundary: annular solve, jumps, QFS sources                                 scalar.py:100-101
  all grid-side sources -> (phys, not in annulus) + interface points               scalar.py:104-105,63-71
  uc[pna] += gu ; helper.correct ; radial -> grid ; mask ; pack                    scalar.py:107-116
"""
import numpy as np
import torch
from ..._lib import lib
from ...device import to_dev, empty, ptr, stream, layer_apply_scalar, HostStaging
from ...embedded_function import EmbeddedFunction, BoundaryFunction
from ...geometry import BoundaryCollection


class ScalarSolver(object):
    kernel_kind = None
    grid_kind = None
    k = 0.0

    def __init__(self, ebdyc, solver_type, helpers, grid_backend, shard_boundaries=False):
        self.ebdyc = ebdyc
        self.solver_type = solver_type
        self.grid_backend = grid_backend
        if helpers is None:
            helpers = [None]*self.ebdyc.N
        self._assign_boundaries(shard_boundaries)
        self.helpers = [self._get_helper(ebdy, helper, owned=(self.owner[b] == self.rank))
                        for b, (ebdy, helper) in enumerate(zip(self.ebdyc, helpers))]
        self.grid = self.ebdyc.grid
        self.kx, self.ky = self.ebdyc.kx, self.ebdyc.ky
        self.ikx, self.iky = self.ebdyc.ikx, self.ebdyc.iky
        self.lap = -self.kx*self.kx - self.ky*self.ky
        self.interpolation_order = 3 if self.solver_type == 'fourth' else np.inf
        self.grid_step = self.ebdyc.grid_step
        self._define_layer_apply()
        self._collect_grid_sources()
        self.dev = self.ebdyc.device
        self._define_grid_evaluator()
        self.iteration_counts = None
        self.timer = None          # set to a device.StageTimer to get per-stage CUDA-event times

    def _assign_boundaries(self, shard):
        """
        Boundary sharding: with `shard_boundaries=True` under torch.distributed every boundary's
        annular solve, QFS applies and correction run on ONE rank (greedy assignment by n_b*M, largest first); the
        grid-sized work is replicated.  Exchanges per solve: a broadcast of every boundary's effective sources
        (S_b doubles) before the evaluation to the grid and of its radial solution (M*n_b doubles) after the
        correction -- both from the owning rank.  Results are bitwise identical on all ranks and to the unsharded run.
        """
        import torch.distributed as dist
        self.world, self.rank = 1, 0
        if shard and dist.is_available() and dist.is_initialized():
            self.world, self.rank = dist.get_world_size(), dist.get_rank()
        cost = [e.bdy.N*e.M for e in self.ebdyc]
        load = [0]*self.world
        self.owner = [0]*len(cost)
        for b in sorted(range(len(cost)), key=lambda i: -cost[i]):
            r = min(range(self.world), key=lambda i: load[i])
            self.owner[b] = r
            load[r] += cost[b]

    def _bcast(self, t, src):
        if self.world > 1:
            import torch.distributed as dist
            dist.broadcast(t, src=src)

    def _collect_grid_sources(self):
        self.grid_sources = BoundaryCollection()
        for helper in self.helpers:
            self.grid_sources.add(helper.interface_qfs_g.source, 'i' if helper.interior else 'e')
        self.grid_sources.amass_information()
        self.d_gsx = to_dev(self.grid_sources.x)
        self.d_gsy = to_dev(self.grid_sources.y)
        self.d_gsw = to_dev(self.grid_sources.weights)
        self.source_offsets = np.concatenate([[0], np.cumsum(self.grid_sources.Ns)])

    def _get_helper(self, ebdy, helper):
        raise NotImplementedError

    # S*T above which the all-pairs sum loses to the spectral-Ewald evaluator (csrc/ewald.cu)
    EWALD_THRESHOLD = 2.0e8
    # spacing of the interface-point evaluator in units of the grid spacing (None: 4 from 4096^2, else 1 = shared box)
    POINT_EVAL_COARSENING = None

    def _define_grid_evaluator(self):
        """
        Reference backends (multi_boundary/modified_helmholtz.py:48-67, poisson.py:40-62): 'pybie2d' / 'fmm2d' /
        'flexmm' are all sums over all sources, a GridBackend object selects the Ewald-like grid evaluator.
        Here: 'direct' = all-pairs device sum (csrc/layer.cu), 'ewald' = spectral-Ewald evaluator
        (csrc/ewald.cu); the reference's names map to 'auto' (by problem size).
        """
        gb = self.grid_backend
        if gb in (None, 'pybie2d', 'fmm2d', 'flexmm', 'auto'):
            work = float(self.grid_sources.N)*float(self.ebdyc.grid_pnai.N)
            gb = 'ewald' if work > self.EWALD_THRESHOLD else 'direct'
        if gb not in ('direct', 'ewald'):
            raise Exception('grid_backend not recognized')
        g = self.ebdyc.grid
        if gb == 'ewald' and abs(g.xh - g.yh) > 1e-14*g.xh:
            gb = 'direct'           # the evaluator needs square cells
        self.evaluation = gb
        self.split_grid_evaluation = gb == 'ewald'
        self._no_interface_field = empty((self.ebdyc.all_ivx.size,))
        self.evaluator_g = None
        self.evaluator_p = None
        if gb == 'ewald':
            from ...device import FreeSpaceEvaluator
            e = self.ebdyc
            pix, piy = np.nonzero(e.phys_not_in_annulus)
            # a single boundary never uses the field on its own interface (helper.correct: it cancels
            # identically), so the second inverse transform + gather is skipped
            pts = (e.all_ivx, e.all_ivy) if e.N > 1 else None
            # Interface points on a coarser Ewald split: on the grid's own spacing the point targets cost a second
            # full-size inverse transform (15 of 41 ms at 16384^2) for 8e4 values.  The split is exact for any
            # spacing; with c*h the FFT box has 1/c^2 of the points and the local sums reach c times further
            # (a few hundred sources per interface point).
            c = self.POINT_EVAL_COARSENING
            if c is None:
                c = 4 if min(g.Nx, g.Ny) >= 4096 else 1
            self.evaluator_p = None
            if pts is not None and c > 1:
                self.evaluator_p = FreeSpaceEvaluator(self.kernel_kind, self.k, float(g.xv[0]), float(g.yv[0]),
                                                      float(g.xh)*c, self.grid_sources.x, self.grid_sources.y,
                                                      points=pts)
                pts = None
            self.evaluator_g = FreeSpaceEvaluator(
                self.kernel_kind, self.k, float(g.xv[0]), float(g.yv[0]), float(g.xh),
                self.grid_sources.x, self.grid_sources.y, grid_nodes=(pix, piy),
                grid_out_pos=e.phys_inds[e.phys_not_in_annulus], points=pts)
        for helper in self.helpers:
            if helper.owned:
                helper.set_evaluation(gb, g, single=self.ebdyc.N == 1)
        self._helper_mode = (gb, self.ebdyc.N == 1)

    def _define_layer_apply(self):
        self.Layer_Apply = self.helpers[0].Layer_Apply

    def get_boundary_values(self, u):
        bv = BoundaryFunction(self.ebdyc)
        bv.load_data([helper.get_boundary_values(ur) for ur, helper in zip(u, self.helpers)])
        return bv

    def get_boundary_normal_derivatives(self, u):
        bv = BoundaryFunction(self.ebdyc)
        bv.load_data([helper.get_boundary_normal_derivatives(ur) for ur, helper in zip(u, self.helpers)])
        return bv

    # ------------------------------------------------------------------------------------
    def solve_device(self, f_dev, **kwargs):
        """f_dev: flat device vector [phys grid | radial blocks] -> flat device vector, same layout."""
        D = self.dev
        L = lib()
        st = stream()
        T = self.timer
        mark = T.mark if T is not None else (lambda name: None)
        mark('start')
        if self.grid_kind == 0 and D.bumpy is None:
            raise Exception('Poisson solver requires ebdyc.ready_bump(...) before solving.')
        L.ipde_grid_scatter_cutoff(D.grid.h, ptr(f_dev), ptr(D.phys_lin), ptr(D.step_compact), D.n_phys,
                                   ptr(D.fc), st)
        L.ipde_grid_solve_scalar(D.grid.h, self.grid_kind, float(self.k), ptr(D.fc), ptr(D.bumpy), ptr(D.uch),
                                 ptr(D.spec_work), ptr(D.uc), st)
        mark('grid_solve')
        # interface values and gradient of the grid solution
        if self.interpolation_order == np.inf:
            nu = D.nufft
            bvs, bxs, bys = nu.apply3(D.uch)
        else:
            ucx, ucy = D.grid.new_real(), D.grid.new_real()
            L.ipde_grid_fd4(D.grid.h, ptr(D.uc), ptr(ucx), ptr(ucy), st)
            bvs, bxs, bys = empty((D.n_iv,)), empty((D.n_iv,)), empty((D.n_iv,))
            for src, dst in ((D.uc, bvs), (ucx, bxs), (ucy, bys)):
                L.ipde_grid_interp_cubic(D.grid.h, ptr(src), D.x0, D.y0, ptr(D.iv_x), ptr(D.iv_y), D.n_iv,
                                         ptr(dst), st)
        self._interface_values = (bvs, bxs, bys)
        mark('interface_interp')
        # helpers may be shared between solvers (helpers=...): make them use this solver's evaluation mode
        for helper in self.helpers:
            helper.evaluation, helper.single = self._helper_mode
        if self.world > 1 and D.bumpy is None and self.grid_kind == 0:
            raise Exception('Poisson solver requires ebdyc.ready_bump(...) before solving.')
        # per-boundary annular solves and effective sources
        sigmag = empty((self.grid_sources.N,))
        its = torch.zeros(len(self.helpers), dtype=torch.int64, device=sigmag.device)
        for b, helper in enumerate(self.helpers):
            if self.owner[b] != self.rank:
                continue
            r0, r1 = D.radial_offsets[b], D.radial_offsets[b + 1]
            o0, o1 = D.bdy_offsets[b], D.bdy_offsets[b + 1]
            fr = f_dev[r0:r1].view(helper.ebdy.M, helper.ebdy.bdy.N)
            sg = helper(fr, bvs[o0:o1], bxs[o0:o1], bys[o0:o1], **kwargs)
            s0, s1 = self.source_offsets[b], self.source_offsets[b + 1]
            L.ipde_vec_mul(ptr(sg), ptr(self.d_gsw[s0:s1]), ptr(sigmag[s0:s1]), int(s1 - s0), st)
            its[b] = helper.iterations_last_call
        if self.world > 1:
            for b in range(len(self.helpers)):
                s0, s1 = self.source_offsets[b], self.source_offsets[b + 1]
                self._bcast(sigmag[s0:s1], self.owner[b])
                self._bcast(its[b:b + 1], self.owner[b])
            self.iteration_counts = [int(v) for v in its.cpu()]
        else:
            self.iteration_counts = [helper.iterations_last_call for helper in self.helpers]
        mark('annular_and_qfs')
        # all grid-side sources -> (phys, not in annulus) U (interface points); pack the grid part of the result
        res = empty((self.ebdyc.dof,))
        L.ipde_grid_gather(ptr(D.uc), ptr(D.phys_lin), D.n_phys, ptr(res), st)
        if self.evaluator_g is not None:
            bu_all = self.evaluator_g(sigmag, grid_out=res)
            if self.evaluator_p is not None:
                bu_all = self.evaluator_p(sigmag)
            if bu_all is None:
                bu_all = self._no_interface_field
        else:
            out = layer_apply_scalar(self.kernel_kind, self.k, self.d_gsx, self.d_gsy, sigmag, D.pnai_x, D.pnai_y)
            L.ipde_index_add(ptr(res), ptr(D.pna_pos), ptr(out), D.n_pna, st)
            bu_all = out[D.n_pna:]
        mark('sources_to_grid')
        # radial part
        for b, helper in enumerate(self.helpers):
            o0, o1 = D.bdy_offsets[b], D.bdy_offsets[b + 1]
            if self.owner[b] == self.rank:
                ur = helper.correct(bu_all[o0:o1])
            else:
                ur = empty((helper.ebdy.M, helper.ebdy.bdy.N))
            self._bcast(ur, self.owner[b])
            mark('correct')
            db = D.ebdys[b]
            db.radial_interp(ur, db.ia_theta, db.ia_t, out=res, out_pos=db.ia_pos)
            r0, r1 = D.radial_offsets[b], D.radial_offsets[b + 1]
            res[r0:r1].copy_(ur.view(-1))
            mark('radial_to_grid')
        return res

    def _staging(self):
        """host staging (device.HostStaging) + device input buffer for the public (host array) entry point"""
        if getattr(self, '_stage', None) is None:
            self._stage = (HostStaging(), empty((self.ebdyc.dof,)))
        return self._stage

    def __call__(self, f, out=None, **kwargs):
        """
        f must be of type EmbeddedFunction; returns a fresh EmbeddedFunction (as the reference does).
        Extension: `out=` an EmbeddedFunction of the same collection to receive the result (a time-stepping caller
        avoids the page faults of a fresh 8*dof-byte host array per call).  EmbeddedFunctions are page-locked, so the
        copies go straight from / to the caller's arrays.
        """
        host, d_in = self._staging()
        host.upload(0, f, d_in)
        res = self.solve_device(d_in, **kwargs)
        if out is None:
            out = EmbeddedFunction(self.ebdyc)
        host.download([(res, out)])
        return out
