"""
ORACLE (test infrastructure, CPU, NumPy) -- never imported by the product.

NumPy restatement of ebdyc.gradient / ebdyc.laplacian (ipde/ebdy_collection.py:708-792) with the per-boundary parts
of ipde/embedded_boundary.py:478-517; the radial -> grid interpolation is the direct NUDFT of oracle/solvers.py.
"""
import numpy as np
from .solvers import interpolate_radial_to_points, fd_x_4, fd_y_4


def _radial_derivatives(e, f):
    ft = np.fft.ifft(np.fft.fft(f)*1j*e.radial_k).real*e.inverse_radial_speed      # embedded_boundary.py:478-481
    fr = e.D00.dot(f)                                                              # :482-483
    b = e.bdy
    return fr*b.normal_x + ft*b.tangent_x, fr*b.normal_y + ft*b.tangent_y          # :484-492


def _radial_laplacian(e, f):
    fx, fy = _radial_derivatives(e, f)
    fxx, _ = _radial_derivatives(e, fx)
    _, fyy = _radial_derivatives(e, fy)
    return fxx + fyy                                                               # :493-498


def _components(ebdyc, ff):
    ff = np.asarray(ff)
    n = ebdyc.grid_phys.N
    f = np.zeros(ebdyc.grid.shape)
    f[ebdyc.phys] = ff[:n]
    return f, [r.copy() for r in ebdyc.v2r(ff[n:])]


def _pack(ebdyc, g, rs):
    for e, r in zip(ebdyc, rs):
        g[e.grid_ia_xind, e.grid_ia_yind] = interpolate_radial_to_points(r, e.grid_ia_r_transf, e.grid_ia_t)
    g = g*ebdyc.phys
    return np.concatenate([g[ebdyc.phys]] + [r.ravel() for r in rs])


def gradient(ebdyc, ff, derivative_type='spectral'):
    f, frs = _components(ebdyc, ff)
    if derivative_type == 'spectral':
        fc = f*ebdyc.grid_step
        fc[ebdyc.ext] = 0.0
        fch = np.fft.fft2(fc)
        fx = np.fft.ifft2(fch*ebdyc.ikx).real
        fy = np.fft.ifft2(fch*ebdyc.iky).real
    else:
        fx, fy = fd_x_4(f, ebdyc.grid.xh), fd_y_4(f, ebdyc.grid.yh)
    rad = [_radial_derivatives(e, fr) for e, fr in zip(ebdyc, frs)]
    return _pack(ebdyc, fx, [r[0] for r in rad]), _pack(ebdyc, fy, [r[1] for r in rad])


def laplacian(ebdyc, ff, derivative_type='spectral'):
    f, frs = _components(ebdyc, ff)
    if derivative_type == 'spectral':
        fc = f*ebdyc.grid_step
        fc[ebdyc.ext] = 0.0
        lapf = np.fft.ifft2(np.fft.fft2(fc)*ebdyc.lap).real
    else:
        fx, fy = fd_x_4(f, ebdyc.grid.xh), fd_y_4(f, ebdyc.grid.yh)
        lapf = fd_x_4(fx, ebdyc.grid.xh) + fd_y_4(fy, ebdyc.grid.yh)
    return _pack(ebdyc, lapf, [_radial_laplacian(e, fr) for e, fr in zip(ebdyc, frs)])
