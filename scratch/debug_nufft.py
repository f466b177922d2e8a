import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import ipde_b200.device as D
from oracle.solvers import nudft2d_type2
for rep in range(2):
  for shape in [(48, 40), (2048, 32), (32, 1024), (40, 8192), (24, 16384)]:
    nx, ny = shape
    xh, yh = 2.4/nx, 3.0/ny
    rng = np.random.default_rng(3)
    f = rng.standard_normal(shape)
    g = D.GridPlan(nx, ny, xh, yh)
    spec = g.fft2(D.to_dev(f))
    x = rng.uniform(0, 2*np.pi, 37)
    y = rng.uniform(0, 2*np.pi, 37)
    x[0], y[0] = 0.0, 0.0
    x[1], y[1] = 2*np.pi*(1 - 1e-12), 1e-9
    nu = D.Nufft2(g, x, y)
    fh = np.fft.fft2(f)
    full = g.spec_to_numpy_full(spec)
    kx = np.fft.fftfreq(nx, xh/(2*np.pi))[:, None]
    ky = np.fft.fftfreq(ny, yh/(2*np.pi))
    ref = nudft2d_type2(np.stack([fh, 1j*kx*fh, 1j*ky*fh]), x, y).real
    scale = [np.abs(f).max(), np.abs(kx).max()*np.abs(f).max(), np.abs(ky).max()*np.abs(f).max()]
    errs = []
    for i, (dx, dy) in enumerate([(0, 0), (1, 0), (0, 1)]):
        v = nu(spec, dx, dy).cpu().numpy()
        e = np.abs(v - ref[i])/scale[i]
        errs.append('%.1e@%d' % (e.max(), e.argmax()))
    for i, v in enumerate(nu.apply3(spec)):
        e = np.abs(v.cpu().numpy() - ref[i])/scale[i]
        errs.append('a3 %.1e@%d' % (e.max(), e.argmax()))
    print(rep, shape, 'fine', nu.fx, nu.fy, 'w', nu.w, 'fft2 err %.1e' % (np.abs(full - fh).max()/np.abs(fh).max()), errs, flush=True)
