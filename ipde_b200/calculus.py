"""
EmbeddedFunction calculus on the device: gradient and laplacian over an EmbeddedBoundaryCollection.

API mirror of ipde/ebdy_collection.py:708-792 (`ebdyc.gradient(ff, derivative_type)`, `ebdyc.laplacian(ff,
derivative_type)`) with the per-boundary parts of ipde/embedded_boundary.py:478-517 (`radial_grid_derivatives`,
`radial_grid_laplacian`, `gradient`, `laplacian`):

  grid     'spectral': cutoff, FFT, i k / -(kx^2+ky^2) symbol, inverse FFT      ebdy_collection.py:727-735,770-774
           'fourth'  : 4th-order centred differences of the uncut values         :736-738
  radial   D00 in r, FFT derivative in t divided by the radial speed, rotation (n, t) -> (x, y)
                                                                                 embedded_boundary.py:478-496
  overwrite the grid values inside every annulus by the radial ones, mask, pack  ebdy_collection.py:739-752

Everything runs through libipde_b200 (the kernels of the solve: scatter+cutoff, the 2-D FFT pair, the derivative
symbols of `k_stokes_stress` -- i k with the self-conjugate Nyquist line dropped, exactly what ifft2(ik .).real does --,
the theta derivative, the radial -> grid interpolation).  SURVEY.md 8(f).3.
"""
import numpy as np
import torch
from ._lib import lib
from .device import to_dev, empty, zeros, ptr, stream, FFT1D


class DeviceCalculus(object):
    def __init__(self, ebdyc):
        self.ebdyc = ebdyc
        self.D = ebdyc.device
        self._b = []
        for e in ebdyc:
            d = lambda a: to_dev(np.ascontiguousarray(a, dtype=float))
            self._b.append(dict(D00=d(e.D00), irs=d(e.inverse_radial_speed), nx=d(e.bdy.normal_x), ny=d(e.bdy.normal_y),
                                tx=d(e.bdy.tangent_x), ty=d(e.bdy.tangent_y), fft=FFT1D(e.bdy.N),
                                work=empty((e.M, e.bdy.N), torch.complex128)))
        self._zero = None

    # ---- radial pieces (device tensors (M, n))
    def _radial_derivatives(self, b, fr):
        """(f_x, f_y) on the radial grid of boundary b"""
        e = self.ebdyc[b]
        B = self._b[b]
        M, n = e.M, e.bdy.N
        L, st = lib(), stream()
        f_r = empty((M, n))
        for i in range(M):                                   # D00 . fr, one row of D00 at a time
            L.ipde_rowvec_apply(ptr(B['D00'][i]), ptr(fr), ptr(f_r[i]), M, n, st)
        f_t = empty((M, n))
        L.ipde_theta_derivative(B['fft'].h, ptr(fr), ptr(B['work']), ptr(f_t), M, st)
        L.ipde_vec_mul(ptr(f_t), ptr(B['irs']), ptr(f_t), M*n, st)
        fx, fy = empty((M, n)), empty((M, n))
        L.ipde_uv_rt(ptr(f_r), ptr(f_t), ptr(B['nx']), ptr(B['ny']), ptr(B['tx']), ptr(B['ty']), M, n, ptr(fx), ptr(fy), 1, st)
        return fx, fy

    def _radial_laplacian(self, b, fr):
        fx, fy = self._radial_derivatives(b, fr)
        fxx, _ = self._radial_derivatives(b, fx)
        _, fyy = self._radial_derivatives(b, fy)
        L = lib()
        L.ipde_vec_axpby(1.0, ptr(fxx), 1.0, ptr(fyy), ptr(fxx), int(fxx.numel()), stream())
        return fxx

    # ---- grid pieces
    def _grid_fields(self, f_dev, kind, derivative_type):
        """full-grid device arrays of the requested derivative(s) of the grid part"""
        D, L, st = self.D, lib(), stream()
        G = D.grid
        n = G.nx*G.ny
        if derivative_type == 'spectral':
            L.ipde_grid_scatter_cutoff(G.h, ptr(f_dev), ptr(D.phys_lin), ptr(D.step_compact), D.n_phys, ptr(D.fc), st)
            spec = G.new_spec()
            L.ipde_grid_fft2_r2c(G.h, ptr(D.fc), ptr(spec), st)
            work = G.new_spec()
            if kind == 'laplacian':
                L.ipde_grid_scalar_multiply(G.h, ptr(spec), 3, 0.0, st)
                out = G.new_real()
                L.ipde_grid_fft2_c2r(G.h, ptr(spec), ptr(work), ptr(out), st)
                return (out,)
            if self._zero is None:
                self._zero = G.new_spec()
                self._zero.zero_()
            sx, sy, sz = G.new_spec(), G.new_spec(), G.new_spec()
            # (2 f_x, f_y, 0) = stress symbols of (u, v, p) = (f, 0, 0)
            L.ipde_grid_stokes_stress(G.h, ptr(spec), ptr(self._zero), ptr(self._zero), ptr(sx), ptr(sy), ptr(sz), st)
            fx, fy = G.new_real(), G.new_real()
            L.ipde_grid_fft2_c2r(G.h, ptr(sx), ptr(work), ptr(fx), st)
            L.ipde_grid_fft2_c2r(G.h, ptr(sy), ptr(work), ptr(fy), st)
            L.ipde_vec_axpby(0.5, ptr(fx), 0.0, ptr(fx), ptr(fx), n, st)
            return fx, fy
        # 'fourth': uncut values on the full grid (zero outside the physical region), centred differences
        one = torch.ones(D.n_phys, dtype=torch.float64, device=f_dev.device)
        L.ipde_grid_scatter_cutoff(G.h, ptr(f_dev), ptr(D.phys_lin), ptr(one), D.n_phys, ptr(D.fc), st)
        fx, fy = G.new_real(), G.new_real()
        L.ipde_grid_fd4(G.h, ptr(D.fc), ptr(fx), ptr(fy), st)
        if kind == 'gradient':
            return fx, fy
        fxx, fyy, tmp = G.new_real(), G.new_real(), G.new_real()
        L.ipde_grid_fd4(G.h, ptr(fx), ptr(fxx), ptr(tmp), st)
        L.ipde_grid_fd4(G.h, ptr(fy), ptr(tmp), ptr(fyy), st)
        L.ipde_vec_axpby(1.0, ptr(fxx), 1.0, ptr(fyy), ptr(fxx), n, st)
        return (fxx,)

    def _pack(self, grid_field, radial_fields):
        """compact [phys grid | radial blocks] vector: grid values, annulus points overwritten by the radial ones"""
        D, L, st = self.D, lib(), stream()
        res = empty((self.ebdyc.dof,))
        L.ipde_grid_gather(ptr(grid_field), ptr(D.phys_lin), D.n_phys, ptr(res), st)
        for b, fr in enumerate(radial_fields):
            db = D.ebdys[b]
            db.radial_interp(fr, db.ia_theta, db.ia_t, out=res, out_pos=db.ia_pos)
            r0, r1 = D.radial_offsets[b], D.radial_offsets[b + 1]
            res[r0:r1].copy_(fr.reshape(-1))
        return res

    def _radial_blocks(self, f_dev):
        D = self.D
        return [f_dev[D.radial_offsets[b]:D.radial_offsets[b + 1]].view(e.M, e.bdy.N) for b, e in enumerate(self.ebdyc)]

    # ---- public (device vectors in the EmbeddedFunction layout)
    def gradient_device(self, f_dev, derivative_type='spectral'):
        gx, gy = self._grid_fields(f_dev, 'gradient', derivative_type)
        rad = [self._radial_derivatives(b, fr) for b, fr in enumerate(self._radial_blocks(f_dev))]
        return self._pack(gx, [r[0] for r in rad]), self._pack(gy, [r[1] for r in rad])

    def laplacian_device(self, f_dev, derivative_type='spectral'):
        (lap,) = self._grid_fields(f_dev, 'laplacian', derivative_type)
        rad = [self._radial_laplacian(b, fr) for b, fr in enumerate(self._radial_blocks(f_dev))]
        return self._pack(lap, rad)
