"""Microbenchmark of the real <-> half-complex ROW transforms (ipde_fft_r2c_rows / ipde_fft_c2r_rows) per row length,
algorithmic GB/s (row of nr reals in, nr/2+1 complex out and back), with cuFFT (torch.fft.rfft / irfft) as the
comparison line.  python scripts/bench_rows.py [nr ...]"""
import sys, os, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
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

reps = int(os.environ.get('REPS', '10'))
for nr in [int(a) for a in sys.argv[1:]] or [4096, 8192, 16384, 32768]:
    batch = (1 << 28)//nr                      # 2.1 GB of reals
    p = D.FFT1D(nr//2)
    f = torch.randn(batch, nr, dtype=torch.float64, device='cuda')
    pitch = nr//2 + 1
    spec = torch.empty(batch, pitch, dtype=torch.complex128, device='cuda')
    out = torch.empty_like(f)
    st = D.stream(); L = lib()
    t_f = timeit(lambda: L.ipde_fft_r2c_rows(p.h, D.ptr(f), D.ptr(spec), batch, nr, pitch, st), reps)
    t_i = timeit(lambda: L.ipde_fft_c2r_rows(p.h, D.ptr(spec), D.ptr(out), batch, pitch, nr, 2.0/nr, st), reps)
    err = float((out - f).abs().max())
    b = batch*(nr*8 + pitch*16)
    line = 'rows nr=%6d batch %6d: r2c %.3f ms %5.0f GB/s | c2r %.3f ms %5.0f GB/s | roundtrip err %.1e' % (
        nr, batch, t_f, b/t_f/1e6, t_i, b/t_i/1e6, err)
    if not os.environ.get('NO_CUFFT'):
        t_cf = timeit(lambda: torch.fft.rfft(f, dim=1), reps)
        sp = torch.fft.rfft(f, dim=1)
        t_ci = timeit(lambda: torch.fft.irfft(sp, n=nr, dim=1), reps)
        del sp
        line += ' | cuFFT %.3f / %.3f ms' % (t_cf, t_ci)
    print(line, flush=True)
    del f, spec, out, p
    torch.cuda.empty_cache()
