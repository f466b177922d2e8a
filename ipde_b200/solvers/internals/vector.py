"""
Per-boundary helper of the vector (Stokes) solvers, device implementation.

API mirror of ipde/solvers/internals/vector.py:8-162 (`VectorHelper`): annular Stokes solve with zero boundary
data in (r, t) components, traction jump across the interface -> QFS effective sources (`sigma_g` for the grid
side, `sigma_r` for the annulus side), and `correct()` which adds the annulus-side stokeslet potential to the
radial solution.  State (`ur`, `vr`, `pr`, `sigma_r`, `sigma_g`) lives on the device between `__call__` and
`correct`, exactly as the reference keeps it on the helper.
"""
import numpy as np
import torch
from ..._lib import lib
from ...annular import ApproximateAnnularGeometry, RealAnnularGeometry
from ...device import to_dev, empty, ptr, stream, gemv, layer_apply_stokes, FFT1D
from .scalar import DeviceQFS, device_qfs, DeviceCurveOperator


def v2f(x):
    return x.reshape(2, x.numel()//2) if torch.is_tensor(x) else x.reshape(2, x.size//2)


class VectorHelper(object):
    def __init__(self, ebdy, annular_solver=None, **kwargs):
        self.ebdy = ebdy
        self.interior = ebdy.interior
        self._extract_extra_kwargs(**kwargs)
        if annular_solver is None:
            self.AAG = ApproximateAnnularGeometry(ebdy.bdy.N, ebdy.M, ebdy.radial_width, ebdy.approximate_radius,
                                                  drop_nyquist=True)
            self._define_annular_solver()
        else:
            self.annular_solver = annular_solver
            self.AAG = self.annular_solver.AAG
        self._set_boundary_estimators()
        self._get_RAG()
        self._get_qfs()
        self._define_layer_apply()
        self._to_device()

    def _extract_extra_kwargs(self, **kwargs):
        pass
    def _define_annular_solver(self):
        raise NotImplementedError
    def _get_qfs(self):
        raise NotImplementedError
    def _define_layer_apply(self):
        raise NotImplementedError

    def _get_RAG(self):
        bb = self.ebdy.bdy if self.interior else self.ebdy.interface
        self.RAG = RealAnnularGeometry(bb.speed, bb.curvature, self.annular_solver.AAG)

    def _set_boundary_estimators(self):
        CO = self.AAG.CO
        if self.interior:
            self._bv_estimator, self._bn_estimator = CO.obc_dirichlet[0], CO.obc_neumann[0]
            self._iv_estimator, self._in_estimator = CO.ibc_dirichlet[0], CO.ibc_neumann[0]
        else:
            self._bv_estimator, self._bn_estimator = CO.ibc_dirichlet[0], CO.ibc_neumann[0]
            self._iv_estimator, self._in_estimator = CO.obc_dirichlet[0], CO.obc_neumann[0]

    def _to_device(self):
        e = self.ebdy
        n, M = e.bdy.N, e.M
        self.d_qfs_g = device_qfs(self.interface_qfs_g)
        self.d_qfs_r = device_qfs(self.interface_qfs_r)
        d = lambda a: to_dev(np.ascontiguousarray(a, dtype=float))
        self.d_bnx, self.d_bny = d(e.bdy.normal_x), d(e.bdy.normal_y)
        self.d_btx, self.d_bty = d(e.bdy.tangent_x), d(e.bdy.tangent_y)
        self.d_inx, self.d_iny = d(e.interface.normal_x), d(e.interface.normal_y)
        self.d_ix, self.d_iy = d(e.interface.x), d(e.interface.y)
        self.d_rx, self.d_ry = d(e.radial_x.ravel()), d(e.radial_y.ravel())
        self.d_D00 = d(e.D00)
        self.d_rsp = d(e.radial_speed)
        self.d_iv_est, self.d_bv_est = d(self._iv_estimator), d(self._bv_estimator)
        self._fft = FFT1D(n)
        self._work = empty((M, n), torch.complex128)
        self._dur = empty((M, n))
        self._tau = empty((4*n,))
        self.curve_w = self.curve_r = None
        if getattr(self.interface_qfs_g, 'structured', False) and getattr(self.interface_qfs_r, 'structured', False):
            from ...qfs_structured import source_to_curves
            self.curve_w = DeviceCurveOperator(source_to_curves(self.interface_qfs_g, [e.interface.c], pressure=True))
            levels = [e.radial_x[i] + 1j*e.radial_y[i] for i in range(M)]
            self.curve_r = DeviceCurveOperator(source_to_curves(self.interface_qfs_r, levels, pressure=True))

    def set_evaluation(self, mode, grid, single):
        """'direct': all-pairs stokeslet sums in correct(); 'ewald': spectral-Ewald evaluators on a box around this
        boundary (annulus-side sources -> own radial points; own grid-side sources -> own interface unless the
        collection has a single boundary, where that term is never used)."""
        self.evaluation = mode
        self.single = single
        if not hasattr(self, 'evaluator_r'):
            self.evaluator_r = None
            self.evaluator_w = None
        if mode == 'ewald' and self.evaluator_r is None and self.curve_r is None:
            from ...device import StokesFreeSpaceEvaluator
            e = self.ebdy
            g, r = self.interface_qfs_g.source, self.interface_qfs_r.source
            x0, y0, h = float(grid.xv[0]), float(grid.yv[0]), float(grid.xh)
            if not single:
                self.evaluator_w = StokesFreeSpaceEvaluator(x0, y0, h, g.x, g.y, points=(e.interface.x, e.interface.y))
            self.evaluator_r = StokesFreeSpaceEvaluator(x0, y0, h, r.x, r.y,
                                                        points=(e.radial_x.ravel(), e.radial_y.ravel()))

    # ---- host conveniences with the reference's names (host arrays in, host arrays out)
    def get_boundary_values(self, fr):
        return self._bv_estimator.dot(fr)
    def get_interface_values(self, fr):
        return self._iv_estimator.dot(fr)
    def get_boundary_traction_uvp(self, u, v, p):
        return self._traction_host(u, v, p, self.d_bv_est)
    def get_interface_traction_uvp(self, u, v, p):
        return self._traction_host(u, v, p, self.d_iv_est)
    def get_boundary_traction_rtp(self, Ur, Ut, p):
        return self.ebdy.convert_uv_to_rt(*self._traction_host(*self.ebdy.convert_rt_to_uv(Ur, Ut), p, self.d_bv_est))
    def get_interface_traction_rtp(self, Ur, Ut, p):
        return self.ebdy.convert_uv_to_rt(*self._traction_host(*self.ebdy.convert_rt_to_uv(Ur, Ut), p, self.d_iv_est))

    def _traction_host(self, u, v, p, est):
        d = lambda a: to_dev(np.ascontiguousarray(a, dtype=float))
        tx, ty = self._traction_uvp(d(u), d(v), d(p), est)
        return tx.cpu().numpy(), ty.cpu().numpy()

    # ---- device
    def _uv_rt(self, a, b, inverse):
        e = self.ebdy
        o0, o1 = torch.empty_like(a), torch.empty_like(b)
        lib().ipde_uv_rt(ptr(a), ptr(b), ptr(self.d_bnx), ptr(self.d_bny), ptr(self.d_btx), ptr(self.d_bty), e.M, e.bdy.N,
                         ptr(o0), ptr(o1), 1 if inverse else 0, stream())
        return o0, o1

    def _traction_uvp(self, u, v, p, est):
        """(Tx, Ty) of the radial field (u, v, p) on the curve picked by `est` (vector.py:87-102)"""
        e = self.ebdy
        n, M = e.bdy.N, e.M
        Ur, Ut = self._uv_rt(u, v, False)
        lib().ipde_theta_derivative(self._fft.h, ptr(Ur), ptr(self._work), ptr(self._dur), M, stream())
        tx, ty = empty((n,)), empty((n,))
        lib().ipde_radial_traction(ptr(Ur), ptr(Ut), ptr(self._dur), ptr(p), ptr(self.d_D00), ptr(est), ptr(self.d_rsp),
                                   ptr(self.d_bnx), ptr(self.d_bny), ptr(self.d_btx), ptr(self.d_bty), M, n, ptr(tx),
                                   ptr(ty), stream())
        return tx, ty

    def __call__(self, fur, fvr, bu, bv, btxx, btxy, btyy, **kwargs):
        e = self.ebdy
        n = e.bdy.N
        fr, ft = self._uv_rt(fur.contiguous(), fvr.contiguous(), False)
        rr, tr, pr = self.annular_solver.solve(self.RAG, fr, ft, None, None, None, None, **kwargs)
        self.iterations_last_call = self.annular_solver.iterations_last_call
        ur, vr = self._uv_rt(rr, tr, True)
        rtx, rty = self._traction_uvp(ur, vr, pr, self.d_iv_est)
        sign = 1.0 if self.interior else -1.0
        lib().ipde_vector_jumps(ptr(rtx), ptr(rty), ptr(btxx), ptr(btxy), ptr(btyy), ptr(bu), ptr(bv), ptr(self.d_inx),
                                ptr(self.d_iny), n, sign, ptr(self._tau), stream())
        self.sigma_g = v2f(self.d_qfs_g(self._tau))
        self.sigma_r = v2f(self.d_qfs_r(self._tau))
        self.ur, self.vr, self.pr = ur, vr, pr
        self._btxx, self._btyy = btxx, btyy
        return self.sigma_g

    def _weighted(self, sigma, w):
        q = torch.empty_like(sigma)
        m = int(w.numel())
        lib().ipde_vec_mul(ptr(sigma[0]), ptr(w), ptr(q[0]), m, stream())
        lib().ipde_vec_mul(ptr(sigma[1]), ptr(w), ptr(q[1]), m, stream())
        return q

    def _own_interface_field(self):
        """(u, v, p) of this boundary's grid-side sources on its own interface"""
        g = self.d_qfs_g
        n = self.ebdy.bdy.N
        if self.curve_w is not None:
            w = empty((3, n))
            self.curve_w(self.sigma_g.reshape(-1), [w[0], w[1], w[2]])
            return w
        q = self._weighted(self.sigma_g, g.sw)
        if getattr(self, 'evaluation', 'direct') == 'ewald' and self.evaluator_w is not None:
            return self.evaluator_w(q[0], q[1])
        return layer_apply_stokes(g.sx, g.sy, q[0], q[1], self.d_ix, self.d_iy)

    def _pressure_continuity(self, pb):
        """shift the radial pressure so that it continues the grid-side pressure across the interface (the two sides'
        stokeslet representations each leave their additive constant free; csrc/vector.cu:k_pressure_continuity)"""
        e = self.ebdy
        lib().ipde_pressure_continuity(ptr(self._btxx), ptr(self._btyy), ptr(pb), ptr(self.d_iv_est), ptr(self.pr), e.M,
                                       e.bdy.N, stream())

    def correct(self, ub, vb, pb, single_ebdy, own_field=None):
        """ub, vb, pb: field of ALL grid-side sources on this boundary's interface (multi-boundary collections; a
        single boundary computes its own)"""
        g, r = self.d_qfs_g, self.d_qfs_r
        n = self.ebdy.bdy.N
        ewald = getattr(self, 'evaluation', 'direct') == 'ewald'
        if single_ebdy:
            sigma_r_tot = self.sigma_r
            w = self._own_interface_field()
            pb = w[2] if w.shape[0] > 2 else None
        else:
            w = own_field if own_field is not None else self._own_interface_field()
            Ub = empty((2*n,))
            lib().ipde_vec_axpby(1.0, ptr(ub), -1.0, ptr(w[0]), ptr(Ub[:n]), n, stream())
            lib().ipde_vec_axpby(1.0, ptr(vb), -1.0, ptr(w[1]), ptr(Ub[n:]), n, stream())
            sigma_r_adj = v2f(r.u2s(Ub))
            sigma_r_tot = torch.empty_like(self.sigma_r)
            lib().ipde_vec_axpby(1.0, ptr(sigma_r_adj), 1.0, ptr(self.sigma_r), ptr(sigma_r_tot),
                                 int(sigma_r_tot.numel()), stream())
        if self.curve_r is not None:
            # structured QFS: stokeslet velocity and pressure on the radial levels as Fourier-banded maps
            rows = []
            for i in range(self.ebdy.M):
                rows += [self.ur[i], self.vr[i], self.pr[i]]
            self.curve_r(sigma_r_tot.reshape(-1), rows, accumulate=True)
        else:
            q = self._weighted(sigma_r_tot, r.sw)
            nt = int(self.d_rx.numel())
            if ewald:
                # accumulated straight into the radial blocks
                self.evaluator_r(q[0], q[1], point_out=(self.ur.view(-1), self.vr.view(-1), self.pr.view(-1)))
            else:
                rslp = layer_apply_stokes(r.sx, r.sy, q[0], q[1], self.d_rx, self.d_ry)
                for dst, k in ((self.ur, 0), (self.vr, 1), (self.pr, 2)):
                    lib().ipde_vec_axpby(1.0, ptr(dst), 1.0, ptr(rslp[k]), ptr(dst), nt, stream())
        if pb is not None:
            self._pressure_continuity(pb.contiguous())
        return self.ur, self.vr, self.pr
