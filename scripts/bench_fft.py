"""Microbenchmark of the FFT kernels: 2-D real FFT pairs of a GridPlan and batched 1-D rows, GB/s of algorithmic
traffic (each pass reads and writes its array once), with cuFFT (torch.fft) as a comparison line."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, numpy as np
import ipde_b200.device as D
from ipde_b200._lib import lib

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps

sizes = [int(a) for a in sys.argv[1:]] or [2048, 4096, 4320, 5120, 8192]
for n in sizes:
    g = D.GridPlan(n, n, 1.0, 1.0)
    f = torch.randn(n, n, dtype=torch.float64, device='cuda')
    spec = g.new_spec(); work = g.new_spec(); out = g.new_real()
    st = D.stream()
    L = lib()
    t_r2c = timeit(lambda: L.ipde_fft_c2c)  # dummy warm
    P = g.pitch
    # individual pieces through the public entry points
    t_f = timeit(lambda: L.ipde_grid_fft2_r2c(g.h, D.ptr(f), D.ptr(spec), st))
    t_i = timeit(lambda: L.ipde_grid_fft2_c2r(g.h, D.ptr(spec), D.ptr(work), D.ptr(out), st))
    real_b, spec_b = n*n*8, n*P*16
    npass_cols = 1 if g.n2 == 1 else 2
    bytes_f = real_b + spec_b + npass_cols*2*spec_b
    print('n=%5d (cols %dx%d): fft2 r2c %.3f ms  %.0f GB/s | c2r %.3f ms %.0f GB/s' %
          (n, g.n1, g.n2, t_f, bytes_f/t_f/1e6, t_i, bytes_f/t_i/1e6), flush=True)
    # comparison line only (north_star): cuFFT through torch.fft on the same arrays -- never used by the library
    t_cf = timeit(lambda: torch.fft.rfft2(f))
    sp = torch.fft.rfft2(f)
    t_ci = timeit(lambda: torch.fft.irfft2(sp, s=(n, n)))
    del sp
    print('        cuFFT (torch.fft) rfft2 %.3f ms | irfft2 %.3f ms   -> ours/cuFFT %.2f / %.2f' %
          (t_cf, t_ci, t_f/t_cf, t_i/t_ci), flush=True)
    # rows only: batched c2c of length n/2
    p = D.FFT1D(n//2)
    x = torch.randn(n, n//2, dtype=torch.complex128, device='cuda')
    y = torch.empty_like(x)
    t = timeit(lambda: p.c2c(x, -1, out=y))
    print('        rows c2c n=%d batch %d: %.3f ms %.0f GB/s' % (n//2, n, t, 2*x.numel()*16/t/1e6), flush=True)
    del g, f, spec, work, out, x, y
    torch.cuda.empty_cache()
