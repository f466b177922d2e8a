"""
Small host helpers with the names the reference exposes in ipde/utilities.py
(affine map :17-34, reversed Chebyshev-Gauss nodes :36-49, Fourier filter :126-162).
The FFT-based helpers of that file (fft2/ifft2 aliases, mfft/mifft,
fourier_multiply) are per-solve work and live in the CUDA library instead.
"""
import numpy as np


def concat(*args):
    return np.concatenate([np.asarray(a).ravel() for a in args])


def affine_transformation(xin, min_in, max_in, min_out, max_out,
                          return_ratio=False, use_numexpr=False):
    """Map [min_in, max_in] affinely onto [min_out, max_out]."""
    rat = (max_out - min_out)/(max_in - min_in)
    xout = (xin - min_in)*rat + min_out
    return (xout, rat) if return_ratio else xout


def get_chebyshev_nodes(lb, ub, order):
    """
    Chebyshev-Gauss nodes, lowest value first, on [-1, 1] and scaled to [lb, ub];
    also returns the scaling ratio (ub-lb)/2.
    """
    xc, _ = np.polynomial.chebyshev.chebgauss(order)
    xc = xc[::-1]
    x, rat = affine_transformation(xc, -1, 1, lb, ub, return_ratio=True)
    return xc, x, rat


class SimpleFourierFilter(object):
    """'fraction' (sharp) or 'rule 36' (exponential) low-pass filter on a mode vector."""
    def __init__(self, modes, filter_type, **kwargs):
        self.n = modes.shape[0]
        self.modes = modes
        self.filter_type = filter_type
        amodes = np.abs(modes)
        max_k = amodes.max()
        if filter_type == 'fraction':
            self.filter = (amodes <= max_k*kwargs['fraction']).astype(float)
        elif filter_type == 'rule 36':
            p = kwargs.get('power', 36)
            self.filter = np.exp(-p*(amodes/max_k)**p)
        else:
            raise Exception('Filter type not defined.')

    def __call__(self, fin, input_type='space', output_type='space'):
        input_is_real = fin.dtype == float and input_type == 'space'
        if input_type == 'space':
            fin = np.fft.fft(fin)
        fout = fin*self.filter
        if output_type == 'space':
            fout = np.fft.ifft(fout)
            if input_is_real:
                fout = fout.real
        return fout
