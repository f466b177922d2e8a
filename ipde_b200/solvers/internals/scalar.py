"""
Per-boundary helper of the scalar solvers (device implementation).

API mirror of ipde/solvers/internals/scalar.py:5-116 (`ScalarHelper`): annular solve with zero
boundary data, interface jumps -> QFS effective sources (`sigma_g` for the grid side, `sigma_r`
for the annulus side), and `correct()` which adds the annulus-side layer potential to the radial
solution.  State (`ur`, `sigma_r`, `sigma_g`) lives on the device between `__call__` and `correct`.
"""
import numpy as np
import torch
from ..._lib import lib
from ...annular import ApproximateAnnularGeometry, RealAnnularGeometry
from ...device import (to_dev, empty, ptr, stream, gemv, layer_apply_scalar, require_cuda)


class DeviceQFS(object):
    """Device copy of the factored operators of an ipde_b200.qfs.QFS object."""
    def __init__(self, qfs):
        self.qfs = qfs
        self.M, self.N = qfs.M, qfs.N
        up = (lambda a: a) if getattr(qfs, 'on_device', False) else to_dev
        if qfs.A is not None:
            self.A, self.Sinv, self.V = up(qfs.A), up(qfs.Sinv), up(qfs.V)
        else:
            self.A = self.Sinv = self.V = None          # built with b2c=False: u2s only
        if qfs.Ab is not None:
            self.Ab, self.Sinv_b, self.Vb = up(qfs.Ab), up(qfs.Sinv_b), up(qfs.Vb)
        else:
            self.Ab = self.Sinv_b = self.Vb = None
        self.sx = to_dev(qfs.source.x)
        self.sy = to_dev(qfs.source.y)
        self.sw = to_dev(qfs.source.weights)
        self.nsrc = qfs.source.N
        nw = max(a.shape[0] for a in (self.A, self.Ab) if a is not None)
        self._w = empty((nw,))

    def _two_step(self, A, Sinv, V, x):
        n = A.shape[0]                      # number of retained singular directions
        w = self._w[:n]
        gemv(A, x, out=w)
        lib().ipde_vec_mul(ptr(w), ptr(Sinv), ptr(w), n, stream())
        return gemv(V, w)

    def __call__(self, tau):
        return self._two_step(self.A, self.Sinv, self.V, tau)

    def u2s(self, ub):
        if self.Ab is None:
            raise Exception('QFS object built without u2s operators')
        return self._two_step(self.Ab, self.Sinv_b, self.Vb, ub)


class DeviceStructuredOperator(object):
    """densities (n_in, N) -> charges (n_out, M): FFT, banded + low-block mat-vec in Fourier space, inverse FFT
    (ipde_b200/qfs_structured.py; kernels k_banded_apply etc. in csrc/dense.cu)"""
    _ffts = {}

    def __init__(self, op):
        self.op = op
        self.N, self.M, self.n_in, self.n_out = op.N, op.M, op.n_in, op.n_out
        # real-to-real operator: output modes k >= 0 only (the kernel mirrors the rest); the arrays stay where the
        # setup built them (the device)
        dev = require_cuda()
        self.band = op.band_t[:, :, :, op.Ks:].to(dev).contiguous()
        self.lowT = op.low_t.transpose(2, 3)[:, :, :, op.KL:].to(dev).contiguous()
        self.npair = (op.n_out + 1)//2
        self._zin = empty((op.n_in, op.N), torch.complex128)
        self._xh = empty((op.n_in, op.N), torch.complex128)
        self._yh = empty((self.npair, op.M), torch.complex128)
        self._yz = empty((self.npair, op.M), torch.complex128)

    @classmethod
    def fft(cls, n):
        from ...device import FFT1D
        if n not in cls._ffts:
            cls._ffts[n] = FFT1D(n)
        return cls._ffts[n]

    def __call__(self, x, out=None):
        L, st, op = lib(), stream(), self.op
        if out is None:
            out = empty((op.n_out*op.M,))
        L.ipde_real_to_complex(ptr(x), ptr(self._zin), op.n_in*op.N, st)
        self.fft(op.N).c2c(self._zin, -1, 1.0, out=self._xh)
        L.ipde_banded_apply_z(ptr(self._xh), op.n_in, op.N, ptr(self.band), ptr(self.lowT), op.n_out, op.BH, op.KL, op.Ks,
                              ptr(self._yh), op.M, st)
        self.fft(op.M).c2c(self._yh, +1, 1.0, out=self._yz)
        o = out.view(op.n_out, op.M)
        for q in range(self.npair):
            im = o[2*q + 1] if 2*q + 1 < op.n_out else None
            L.ipde_complex_parts(ptr(self._yz[q]), ptr(o[2*q]), ptr(im), op.M, st)
        return out


class DeviceStructuredQFS(object):
    """device side of ipde_b200.qfs_structured.StructuredQFS; same surface as DeviceQFS"""
    def __init__(self, qfs):
        self.qfs = qfs
        self.M, self.N = qfs.M, qfs.N
        self.op = DeviceStructuredOperator(qfs.op) if qfs.op is not None else None
        self.op_u2s = DeviceStructuredOperator(qfs.op_u2s) if qfs.op_u2s is not None else None
        self.sx = to_dev(qfs.source.x)
        self.sy = to_dev(qfs.source.y)
        self.sw = to_dev(qfs.source.weights)
        self.nsrc = qfs.source.N

    def __call__(self, tau):
        return self.op(tau)

    def u2s(self, ub):
        if self.op_u2s is None:
            raise Exception('QFS object built without u2s operators')
        return self.op_u2s(ub)


class DeviceCurveOperator(object):
    """
    charges on a StructuredQFS source curve -> field on curves sampled at the boundary parameters (own interface,
    radial levels): FFT of the charges, banded + low-block mat-vec (k_banded_apply), inverse FFTs on the source grid,
    every rho-th sample kept.  Replaces the per-boundary sums of ScalarHelper.correct / VectorHelper.correct
    (ipde/solvers/internals/scalar.py:101,113; vector.py:141,156) for boundaries with structured QFS operators.
    """
    def __init__(self, op):
        self.op = op
        dev = require_cuda()
        self.band = op.band_half_t.to(dev).contiguous()
        self.lowT = op.low_t.transpose(2, 3)[:, :, :, op.KL:].to(dev).contiguous()
        self.npair = (op.n_out + 1)//2
        self._zin = empty((op.n_in, op.M), torch.complex128)
        self._xh = empty((op.n_in, op.M), torch.complex128)
        self._yh = empty((self.npair, op.M), torch.complex128)
        self._yz = empty((self.npair, op.M), torch.complex128)

    def __call__(self, q, out, accumulate=False):
        """q: (n_in*M,) charges; out: (n_out, N) rows (any row-contiguous device views, given as a list or a tensor)"""
        L, st, op = lib(), stream(), self.op
        L.ipde_real_to_complex(ptr(q), ptr(self._zin), op.n_in*op.M, st)
        DeviceStructuredOperator.fft(op.M).c2c(self._zin, -1, 1.0, out=self._xh)
        L.ipde_banded_apply_z(ptr(self._xh), op.n_in, op.M, ptr(self.band), ptr(self.lowT), op.n_out, op.BH, op.KL, op.Ks,
                              ptr(self._yh), op.M, st)
        DeviceStructuredOperator.fft(op.M).c2c(self._yh, +1, 1.0, out=self._yz)
        rho = op.M//op.N
        for qp in range(self.npair):
            re = out[2*qp]
            im = out[2*qp + 1] if 2*qp + 1 < op.n_out else None
            L.ipde_complex_parts_strided(ptr(self._yz[qp]), ptr(re), ptr(im), op.N, rho, 1 if accumulate else 0, st)
        return out


def device_qfs(qfs):
    return DeviceStructuredQFS(qfs) if getattr(qfs, 'structured', False) else DeviceQFS(qfs)


class ScalarHelper(object):
    kernel_kind = None       # 0 Laplace, 1 modified Helmholtz (csrc/layer.cu)
    k = 0.0

    def __init__(self, ebdy, helper=None, grid_backend='pybie2d', owned=True):
        self.ebdy = ebdy
        self.grid_backend = grid_backend
        self.interior = ebdy.interior
        # owned=False: this boundary's annular solve / QFS / correction run on another rank (boundary sharding);
        # only the source geometry is built here
        self.owned = owned
        if not owned:
            self._get_qfs()
            self._define_layer_apply()
            self.iterations_last_call = -1
            return
        if helper is None:
            self.AAG = ApproximateAnnularGeometry(ebdy.bdy.N, ebdy.M, ebdy.radial_width, ebdy.approximate_radius)
            self._define_annular_solver()
        else:
            self.annular_solver = helper.annular_solver if hasattr(helper, 'annular_solver') else helper
            self.AAG = self.annular_solver.AAG
        self._set_boundary_estimators()
        self._get_RAG()
        self._get_qfs()
        self._define_layer_apply()
        self._to_device()

    def _define_annular_solver(self):
        raise NotImplementedError
    def _get_qfs(self):
        raise NotImplementedError

    def _get_RAG(self):
        bb = self.ebdy.bdy if self.interior else self.ebdy.interface
        self.RAG = RealAnnularGeometry(bb.speed, bb.curvature, self.annular_solver.AAG)

    def _set_boundary_estimators(self):
        CO = self.annular_solver.AAG.CO
        if self.interior:
            self._bv_estimator, self._bn_estimator = CO.obc_dirichlet[0], CO.obc_neumann[0]
            self._iv_estimator, self._in_estimator = CO.ibc_dirichlet[0], CO.ibc_neumann[0]
        else:
            self._bv_estimator, self._bn_estimator = CO.ibc_dirichlet[0], CO.ibc_neumann[0]
            self._iv_estimator, self._in_estimator = CO.obc_dirichlet[0], CO.obc_neumann[0]

    def _to_device(self):
        e = self.ebdy
        self.d_in_estimator = to_dev(self._in_estimator)
        self.d_qfs_g = device_qfs(self.interface_qfs_g)
        self.d_qfs_r = device_qfs(self.interface_qfs_r)
        self.d_inx = to_dev(e.interface.normal_x)
        self.d_iny = to_dev(e.interface.normal_y)
        self.d_ix = to_dev(e.interface.x)
        self.d_iy = to_dev(e.interface.y)
        self.d_rx = to_dev(e.radial_x.ravel())
        self.d_ry = to_dev(e.radial_y.ravel())
        n = e.bdy.N
        self._tau = empty((2*n,))
        self._w = empty((n,))
        self.curve_w = self.curve_r = None
        if getattr(self.interface_qfs_g, 'structured', False) and getattr(self.interface_qfs_r, 'structured', False):
            from ...qfs_structured import source_to_curves
            self.curve_w = DeviceCurveOperator(source_to_curves(self.interface_qfs_g, [e.interface.c]))
            levels = [e.radial_x[i] + 1j*e.radial_y[i] for i in range(e.M)]
            self.curve_r = DeviceCurveOperator(source_to_curves(self.interface_qfs_r, levels))

    def set_evaluation(self, mode, grid, single):
        """
        mode 'direct': all-pairs sums in correct(); 'ewald': spectral-Ewald evaluators on a box around this
        boundary, aligned with the background grid (own grid-side sources -> own interface, annulus-side sources ->
        own radial points).  `single`: the collection has one boundary, so the field of the other boundaries on
        this interface vanishes identically and the first evaluation is skipped (cf. the reference's own
        `single_ebdy` shortcut in VectorHelper.correct, ipde/solvers/internals/vector.py:138-139).
        """
        self.evaluation = mode
        self.single = single
        if not hasattr(self, 'evaluator_w'):
            self.evaluator_w = None
            self.evaluator_r = None
        if mode == 'ewald' and self.curve_r is None:
            from ...device import FreeSpaceEvaluator
            e = self.ebdy
            g, r = self.interface_qfs_g.source, self.interface_qfs_r.source
            x0, y0, h = float(grid.xv[0]), float(grid.yv[0]), float(grid.xh)
            # a helper first used in a single-boundary collection and later reused in a larger one still needs the
            # own-interface evaluator: created whenever it is missing, independent of evaluator_r
            if not single and self.evaluator_w is None:
                self.evaluator_w = FreeSpaceEvaluator(self.kernel_kind, self.k, x0, y0, h, g.x, g.y,
                                                      points=(e.interface.x, e.interface.y))
            if self.evaluator_r is None:
                self.evaluator_r = FreeSpaceEvaluator(self.kernel_kind, self.k, x0, y0, h, r.x, r.y,
                                                      points=(e.radial_x.ravel(), e.radial_y.ravel()))

    # ---- small (1 x M) . (M, n) products; host arrays in, host arrays out (drop-in use)
    def get_boundary_values(self, ur):
        return self._bv_estimator.dot(ur)
    def get_boundary_normal_derivatives(self, ur):
        return self._bn_estimator.dot(ur)
    def get_interface_values(self, ur):
        return self._iv_estimator.dot(ur)
    def get_interface_normal_derivatives(self, ur):
        return self._in_estimator.dot(ur)

    def _define_layer_apply(self):
        kind, k = self.kernel_kind, self.k
        def Layer_Apply(src, trg, ch):
            """u(trg) = sum G(trg - src) ch*src.weights, host arrays (reference Layer_Apply signature)"""
            q = to_dev(np.asarray(ch, dtype=float)*src.weights)
            out = layer_apply_scalar(kind, k, to_dev(src.x), to_dev(src.y), q, to_dev(trg.x), to_dev(trg.y))
            return out.cpu().numpy()
        self.Layer_Apply = Layer_Apply

    # ---- per-solve, device tensors in and out
    def __call__(self, fr, bv, bx, by, **kwargs):
        e = self.ebdy
        n, M = e.bdy.N, e.M
        ur = self.annular_solver.solve(self.RAG, fr, None, None, **kwargs)
        self.iterations_last_call = self.annular_solver.iterations_last_call
        sign = 1.0 if self.interior else -1.0
        lib().ipde_scalar_jumps(ptr(bv), ptr(bx), ptr(by), ptr(self.d_inx), ptr(self.d_iny), ptr(ur),
                                ptr(self.d_in_estimator), M, n, sign, ptr(self._tau), stream())
        self.sigma_g = self.d_qfs_g(self._tau)
        self.sigma_r = self.d_qfs_r(self._tau)
        self.ur = ur
        return self.sigma_g

    def correct(self, ub):
        g, r = self.d_qfs_g, self.d_qfs_r
        n = self.ebdy.bdy.N
        ewald = getattr(self, 'evaluation', 'direct') == 'ewald'
        qr = empty((r.nsrc,))
        if self.curve_r is not None:
            # structured QFS: own-source fields on the own interface and the radial levels as Fourier-banded maps
            # (source weights are 1: the strengths are charges)
            if ewald and self.single:
                qr = self.sigma_r
            else:
                self.curve_w(self.sigma_g, [self._w])
                lib().ipde_vec_axpby(1.0, ptr(ub), -1.0, ptr(self._w), ptr(self._w), n, stream())
                sigma_r_adj = r.u2s(self._w)
                lib().ipde_vec_axpby(1.0, ptr(sigma_r_adj), 1.0, ptr(self.sigma_r), ptr(qr), r.nsrc, stream())
            self.curve_r(qr, self.ur, accumulate=True)
            return self.ur
        if ewald and self.single:
            # ub is the field of our own sources only: ub - w == 0 and the adjustment vanishes
            lib().ipde_vec_mul(ptr(self.sigma_r), ptr(r.sw), ptr(qr), r.nsrc, stream())
        else:
            # field of our own grid-side sources on our own interface
            q = empty((g.nsrc,))
            lib().ipde_vec_mul(ptr(self.sigma_g), ptr(g.sw), ptr(q), g.nsrc, stream())
            if ewald:
                self.evaluator_w(q, point_out=self._w)
            else:
                layer_apply_scalar(self.kernel_kind, self.k, g.sx, g.sy, q, self.d_ix, self.d_iy, out=self._w)
            lib().ipde_vec_axpby(1.0, ptr(ub), -1.0, ptr(self._w), ptr(self._w), n, stream())
            sigma_r_adj = r.u2s(self._w)
            lib().ipde_vec_axpby(1.0, ptr(sigma_r_adj), 1.0, ptr(self.sigma_r), ptr(qr), r.nsrc, stream())
            lib().ipde_vec_mul(ptr(qr), ptr(r.sw), ptr(qr), r.nsrc, stream())
        if ewald:
            self.evaluator_r(qr, point_out=self.ur.view(-1), accumulate_points=True)
        else:
            layer_apply_scalar(self.kernel_kind, self.k, r.sx, r.sy, qr, self.d_rx, self.d_ry,
                               out=self.ur.view(-1), accumulate=True)
        return self.ur
