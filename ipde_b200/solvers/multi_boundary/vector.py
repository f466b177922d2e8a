"""
VectorSolver: the inhomogeneous vector (Stokes) solve over an EmbeddedBoundaryCollection, on the device.

API mirror of ipde/solvers/multi_boundary/vector.py:6-112.  `__call__(fu, fv, **kwargs)` takes two
EmbeddedFunctions (host) and returns three (u, v, p); kwargs (tol, restart, maxiter, verbose) go to the annular
GMRES.  On the device:

  scatter + cutoff of (fu, fv) -> demean -> 2 forward FFTs -> Stokes symbol -> 3 inverse FFTs          vector.py:59-63
  interface values of u, v and of the grid stress: spectral mode forms the stress spectra txx = 2 u_x - p,
     txy = u_y + v_x, tyy = 2 v_y - p directly from the spectra of (u, v, p) (one elementwise kernel; the
     reference goes through 8 more grid FFTs, vector.py:68-73) and interpolates the five fields by type-2 NUFFT;
     'fourth': 4th-order differences + cubic interpolation of the five grid fields, as the reference.
  per boundary: annular Stokes solve, traction jumps, QFS sources                                      vector.py:82-84
  all grid-side stokeslets -> (phys, not in annulus) + interface points: (u, v, p)                     vector.py:86-95
  helper.correct ; radial -> grid ; mask ; pack                                                       vector.py:97-111
"""
import numpy as np
import torch
from ..._lib import lib
from ...device import to_dev, empty, ptr, stream, layer_apply_stokes, HostStaging
from ...embedded_function import EmbeddedFunction, BoundaryFunction
from ...geometry import BoundaryCollection


class VectorSolver(object):
    def __init__(self, ebdyc, solver_type='spectral', helpers=None, **kwargs):
        self.ebdyc = ebdyc
        self.solver_type = solver_type
        if helpers is None:
            helpers = [None]*self.ebdyc.N
        self._extract_extra_kwargs(**kwargs)
        self.helpers = [self._get_helper(ebdy, helper) for ebdy, helper in zip(self.ebdyc.ebdys, helpers)]
        self.AS_list = [helper.annular_solver for helper in self.helpers]
        self.grid = self.ebdyc.grid
        self.kx, self.ky = self.ebdyc.kx, self.ebdyc.ky
        self.ikx, self.iky = self.ebdyc.ikx, self.ebdyc.iky
        self.lap = -self.kx*self.kx - self.ky*self.ky
        self._get_specific_operators()
        self.interpolation_order = 3 if self.solver_type == 'fourth' else np.inf
        self.grid_step = self.ebdyc.grid_step
        self._define_layer_apply()
        self._collect_grid_sources()
        self.dev = self.ebdyc.device
        self.iteration_counts = None
        self._define_grid_evaluator()
        self._assign_boundaries(bool(kwargs.get('shard_boundaries', False)))

    def _assign_boundaries(self, shard):
        """
        Boundary sharding, per-solve work only: with `shard_boundaries=True` under torch.distributed
        every boundary's annular Stokes solve, QFS applies and correction run on ONE rank (greedy by n_b*M, largest
        first); the grid-sized work is replicated.  Exchanges per solve: a broadcast of every boundary's effective
        sources (2 S_b doubles) and of its radial solution (3 M n_b doubles) from the owner.  Unlike the scalar
        solver, all ranks still build every helper (setup is replicated).
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

    # S*T above which the all-pairs stokeslet sum loses to the spectral-Ewald evaluator
    EWALD_THRESHOLD = 2.0e8

    def _define_grid_evaluator(self):
        """'direct' = all-pairs stokeslet sums (csrc/layer.cu), 'ewald' = StokesFreeSpaceEvaluator (three Laplace
        spectral-Ewald passes with gradient fields); None / 'auto' chooses by problem size."""
        gb = getattr(self, 'grid_backend', None)
        if gb in (None, 'auto', 'pybie2d'):
            work = float(self.grid_sources.N)*float(self.ebdyc.grid_pnai.N)
            gb = 'ewald' if work > self.EWALD_THRESHOLD else 'direct'
        if gb not in ('direct', 'ewald'):
            raise Exception('grid_backend not recognized')
        g = self.ebdyc.grid
        if gb == 'ewald' and abs(g.xh - g.yh) > 1e-14*g.xh:
            gb = 'direct'
        self.evaluation = gb
        self.evaluator_g = None
        e = self.ebdyc
        single = e.N == 1
        if gb == 'ewald':
            from ...device import StokesFreeSpaceEvaluator
            pix, piy = np.nonzero(e.phys_not_in_annulus)
            pts = (e.all_ivx, e.all_ivy) if not single else None
            self.evaluator_g = StokesFreeSpaceEvaluator(float(g.xv[0]), float(g.yv[0]), float(g.xh), self.grid_sources.x,
                                                        self.grid_sources.y, grid_nodes=(pix, piy),
                                                        grid_out_pos=e.phys_inds[e.phys_not_in_annulus], points=pts)
        for helper in self.helpers:
            helper.set_evaluation(gb, g, single)
        self._helper_mode = (gb, single)

    def _collect_grid_sources(self):
        self.grid_sources = BoundaryCollection()
        for helper in self.helpers:
            self.grid_sources.add(helper.interface_qfs_g.source, 'i' if helper.interior else 'e')
        self.grid_sources.amass_information()
        self.d_gsx = to_dev(self.grid_sources.x)
        self.d_gsy = to_dev(self.grid_sources.y)
        self.d_gsw = to_dev(self.grid_sources.weights)
        self.source_offsets = np.concatenate([[0], np.cumsum(self.grid_sources.Ns)])

    def _extract_extra_kwargs(self, **kwargs):
        self.grid_backend = kwargs.get('grid_backend', None)
    def _get_helper(self, ebdy, AS):
        raise NotImplementedError
    def _get_specific_operators(self):
        pass
    def _define_layer_apply(self):
        self.Layer_Apply = self.helpers[0].Layer_Apply

    def get_boundary_values(self, frs):
        return np.concatenate([helper.get_boundary_values(fr) for fr, helper in zip(frs, self.helpers)])

    def get_boundary_tractions(self, urs, vrs, prs):
        return np.concatenate([np.concatenate(helper.get_boundary_traction_uvp(ur, vr, pr))
                               for ur, vr, pr, helper in zip(urs, vrs, prs, self.helpers)])

    # ------------------------------------------------------------------------------------
    def solve_device(self, fu_dev, fv_dev, **kwargs):
        """flat device vectors [phys grid | radial blocks] of both forcing components -> (u, v, p) likewise"""
        D = self.dev
        L = lib()
        st = stream()
        if D.bumpy is None:
            raise Exception('Stokes solver requires embedded boundary collection with a bump function.')
        G = D.grid
        if getattr(self, '_vbuf', None) is None:
            self._vbuf = dict(fvc=G.new_real(), vc=G.new_real(), pc=G.new_real(), vh=G.new_spec(), ph=G.new_spec())
        B = self._vbuf
        L.ipde_grid_scatter_cutoff(G.h, ptr(fu_dev), ptr(D.phys_lin), ptr(D.step_compact), D.n_phys, ptr(D.fc), st)
        L.ipde_grid_scatter_cutoff(G.h, ptr(fv_dev), ptr(D.phys_lin), ptr(D.step_compact), D.n_phys, ptr(B['fvc']), st)
        L.ipde_grid_solve_stokes(G.h, ptr(D.fc), ptr(B['fvc']), ptr(D.bumpy), ptr(D.uch), ptr(B['vh']), ptr(B['ph']),
                                 ptr(D.spec_work), ptr(D.uc), ptr(B['vc']), ptr(B['pc']), st)
        uc, vc, pc = D.uc, B['vc'], B['pc']
        n_iv = D.n_iv
        if self.interpolation_order == np.inf:
            # Five fine grids, as the reference interpolates (u, v, txx, txy, tyy): the stress is formed in spectral
            # space with the Nyquist row/column of the derivative symbols dropped, exactly what `fourier` +
            # ifft2(.).real do (vector.py:68-77).  Taking the gradient of the interpolants of u and v instead (three
            # fine grids) differs by the derivative of the Nyquist modes and was measured to break the 1e-11 parity at
            # the resolutions of the tests.
            nu = D.nufft
            if 'txx' not in B:
                B.update(txx=G.new_spec(), txy=G.new_spec(), tyy=G.new_spec())
            L.ipde_grid_stokes_stress(G.h, ptr(D.uch), ptr(B['vh']), ptr(B['ph']), ptr(B['txx']), ptr(B['txy']),
                                      ptr(B['tyy']), st)
            bus, bvs = nu(D.uch, 0, 0), nu(B['vh'], 0, 0)
            btxx, btxy, btyy = nu(B['txx'], 0, 0), nu(B['txy'], 0, 0), nu(B['tyy'], 0, 0)
        else:
            ucx, ucy, vcx, vcy = G.new_real(), G.new_real(), G.new_real(), G.new_real()
            L.ipde_grid_fd4(G.h, ptr(uc), ptr(ucx), ptr(ucy), st)
            L.ipde_grid_fd4(G.h, ptr(vc), ptr(vcx), ptr(vcy), st)
            nn = G.nx*G.ny
            L.ipde_vec_axpby(2.0, ptr(ucx), -1.0, ptr(pc), ptr(ucx), nn, st)      # tcxx
            L.ipde_vec_axpby(1.0, ptr(ucy), 1.0, ptr(vcx), ptr(ucy), nn, st)       # tcxy
            L.ipde_vec_axpby(2.0, ptr(vcy), -1.0, ptr(pc), ptr(vcy), nn, st)      # tcyy
            outs = [empty((n_iv,)) for _ in range(5)]
            for src, dst in zip((uc, vc, ucx, ucy, vcy), outs):
                L.ipde_grid_interp_cubic(G.h, ptr(src), D.x0, D.y0, ptr(D.iv_x), ptr(D.iv_y), n_iv, ptr(dst), st)
            bus, bvs, btxx, btxy, btyy = outs
        for helper in self.helpers:
            helper.evaluation, helper.single = self._helper_mode
        # per-boundary annular solves and effective sources
        nsrc = self.grid_sources.N
        sigmag = empty((2, nsrc))
        its = torch.zeros(len(self.helpers), dtype=torch.int64, device=sigmag.device)
        for b, helper in enumerate(self.helpers):
            if self.owner[b] != self.rank:
                continue
            r0, r1 = D.radial_offsets[b], D.radial_offsets[b + 1]
            o0, o1 = D.bdy_offsets[b], D.bdy_offsets[b + 1]
            M, n = helper.ebdy.M, helper.ebdy.bdy.N
            sg = helper(fu_dev[r0:r1].view(M, n), fv_dev[r0:r1].view(M, n), bus[o0:o1], bvs[o0:o1], btxx[o0:o1],
                        btxy[o0:o1], btyy[o0:o1], **kwargs)
            s0, s1 = self.source_offsets[b], self.source_offsets[b + 1]
            for c in range(2):
                L.ipde_vec_mul(ptr(sg[c]), ptr(self.d_gsw[s0:s1]), ptr(sigmag[c, s0:s1]), int(s1 - s0), st)
            its[b] = helper.iterations_last_call
        if self.world > 1:
            for b in range(len(self.helpers)):
                s0, s1 = self.source_offsets[b], self.source_offsets[b + 1]
                for c in range(2):
                    self._bcast(sigmag[c, s0:s1], self.owner[b])
                self._bcast(its[b:b + 1], self.owner[b])
            self.iteration_counts = [int(v) for v in its.cpu()]
        else:
            self.iteration_counts = [helper.iterations_last_call for helper in self.helpers]
        # all grid-side stokeslets -> (phys, not in annulus) U (interface points)
        dof = self.ebdyc.dof
        res = [empty((dof,)) for _ in range(3)]
        for k, field in enumerate((uc, vc, pc)):
            L.ipde_grid_gather(ptr(field), ptr(D.phys_lin), D.n_phys, ptr(res[k]), st)
        single = len(self.ebdyc) == 1
        if self.evaluator_g is not None:
            iface = self.evaluator_g(sigmag[0].contiguous(), sigmag[1].contiguous(), grid_out=(res[0], res[1], res[2]))
            if iface is None:
                iface = empty((3, n_iv))           # single boundary: never read
        else:
            out = layer_apply_stokes(self.d_gsx, self.d_gsy, sigmag[0], sigmag[1], D.pnai_x, D.pnai_y)
            for k in range(3):
                L.ipde_index_add(ptr(res[k]), ptr(D.pna_pos), ptr(out[k]), D.n_pna, st)
            iface = out[:, D.n_pna:]
        for b, helper in enumerate(self.helpers):
            o0, o1 = D.bdy_offsets[b], D.bdy_offsets[b + 1]
            if self.owner[b] == self.rank:
                bi = [iface[k, o0:o1].contiguous() for k in range(3)]
                urs = [u.contiguous() for u in helper.correct(bi[0], bi[1], bi[2], single)]
            else:
                urs = [empty((helper.ebdy.M, helper.ebdy.bdy.N)) for _ in range(3)]
            for k in range(3):
                self._bcast(urs[k], self.owner[b])
            db = D.ebdys[b]
            r0, r1 = D.radial_offsets[b], D.radial_offsets[b + 1]
            for k in range(3):
                db.radial_interp(urs[k], db.ia_theta, db.ia_t, out=res[k], out_pos=db.ia_pos)
                res[k][r0:r1].copy_(urs[k].reshape(-1))
        return res

    def _staging(self):
        """host staging (device.HostStaging) + device input buffers for the public (host array) entry point"""
        if getattr(self, '_stage', None) is None:
            n = self.ebdyc.dof
            self._stage = (HostStaging(), [empty((n,)), empty((n,))])
        return self._stage

    def __call__(self, fu, fv, out=None, **kwargs):
        """
        fu, fv: EmbeddedFunctions -> (u, v, p) fresh EmbeddedFunctions, as ipde/solvers/multi_boundary/vector.py:57-112.
        Extension: `out=(u, v, p)` EmbeddedFunctions of the same collection to receive the result (time stepping).
        """
        host, d_in = self._staging()
        for key, (src, db) in enumerate(zip((fu, fv), d_in)):
            host.upload(key, src, db)
        res = self.solve_device(d_in[0], d_in[1], **kwargs)
        if out is None:
            out = tuple(EmbeddedFunction(self.ebdyc) for _ in range(3))
        host.download(list(zip(res, out)))
        return tuple(out)
