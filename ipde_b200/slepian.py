"""
Cutoff ("step") and bump functions built from a discrete prolate spheroidal
(Slepian) window.

Mirrors `SlepianMollifier` of the reference (ipde/slepian/chebeval_bump_step.py:23-44),
which evaluates *precomputed* Chebyshev coefficient tables
(ipde/slepian/heaviside_coefficients.py) by Clenshaw recurrences
(ipde/slepian/chebeval.py:15-29).  Those tables were produced by
ipde/slepian/construct_coefficients.py from
ipde/slepian/function_generator_bump_step.py:34-45: a dpss(1000, r/4) window,
quintic-spline interpolated, integrated and normalised to give the step, then
fit by an even Chebyshev series whose order grows until the last coefficient
is below 1e-14.  We regenerate the coefficients with the same recipe at
construction time (SciPy only; the spline is integrated exactly instead of by
adaptive quadrature), so any r -- including the non-integer r=1.5*M that the
examples ask for and the tables cannot serve -- is available.  Agreement with
the reference tables for integer r is checked in tests/test_slepian.py against
values frozen in tests/golden/slepian_golden.npz.

This is setup-time code: cutoffs are evaluated once at grid registration; the
per-solve path only sees the resulting `grid_step` array.
"""
import numpy as np
import scipy.interpolate
import scipy.signal

_cache = {}


def _fit_even(f, order):
    x, _ = np.polynomial.chebyshev.chebgauss(order)
    c = np.polynomial.chebyshev.chebfit(x, f(x), order - 1)
    return c[::2]


def _build(r, N=1000):
    x = np.linspace(-1.0, 1.0, N, endpoint=True)
    w = scipy.signal.windows.dpss(N, 0.25*float(r))
    bump = scipy.interpolate.InterpolatedUnivariateSpline(x, w, k=5)
    anti = bump.antiderivative()
    w_int = anti(x) - anti(-1.0)
    w_int /= w_int[-1]
    step = scipy.interpolate.InterpolatedUnivariateSpline(x, w_int, k=5)

    def odd_part(z):
        z = np.asarray(z, dtype=float)
        out = np.empty_like(z)
        small = np.abs(z) < 1e-8
        zs = np.where(small, 1.0, z)
        out[:] = (step(zs) - 0.5)/zs
        if small.any():
            out[small] = step.derivative()(0.0)
        return out

    def fit(f):
        order = 8
        while True:
            c = _fit_even(f, order)
            if not (np.abs(c[-1]) > 1e-14) or order >= 400:
                return c
            order += 2
    return fit(bump), fit(odd_part)


def _clenshaw(x, c):
    # same recurrence as the reference's numba_chebeval1 (chebeval.py:15-23)
    x2 = 2.0*x
    if len(c) == 1:
        return np.full_like(x, c[0])
    c0 = np.full_like(x, c[-2])
    c1 = np.full_like(x, c[-1])
    for i in range(3, len(c) + 1):
        tmp = c0
        c0 = c[-i] - c1
        c1 = tmp + c1*x2
    return c0 + c1*x


class SlepianMollifier(object):
    """
    step(x): 0 for x <= -1, 1 for x >= 1, smooth monotone in between.
    bump(x): even, supported on [-1, 1], max 1 at 0.
    """
    def __init__(self, r):
        self.r = r
        key = float(r)
        if key not in _cache:
            _cache[key] = _build(key)
        self.bump_c, self.step_c = _cache[key]

    def bump(self, x, check_bounds=True):
        scalar = np.isscalar(x)
        x = np.atleast_1d(np.asarray(x, dtype=float))
        out = _clenshaw(2.0*x*x - 1.0, self.bump_c)
        if check_bounds:
            out = np.where(np.logical_and(x >= -1.0, x <= 1.0), out, 0.0)
        return float(out[0]) if scalar else out

    def step(self, x, check_bounds=True):
        scalar = np.isscalar(x)
        x = np.atleast_1d(np.asarray(x, dtype=float))
        out = 0.5 + x*_clenshaw(2.0*x*x - 1.0, self.step_c)
        if check_bounds:
            out = np.where(x < -1.0, 0.0, out)
            out = np.where(x > 1.0, 1.0, out)
        return float(out[0]) if scalar else out
