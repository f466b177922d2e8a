"""
EmbeddedFunction / BoundaryFunction: flat ndarray subclasses holding
[physical grid values | radial block of ebdy 0 | radial block of ebdy 1 | ...]
and [boundary values of ebdy 0 | ebdy 1 | ...].

API mirror of ipde/embedded_function.py:16-342 (constructor keywords :24, layout
:62-79, `f[i]` / `f['grid']` accessors :84-101, `load_data` :105-113,
`define_via_function` :126-132, `get_components` :135-138, `get_grid_value` /
`get_smoothed_grid_value` :184-194).  These are host containers -- the solvers
copy the flat vector to the device and do the scatter/cutoff/gather
(reference :184-194 and :105-113) there, fused into the first and last grid
kernels; the host methods below exist for drop-in use by callers.
"""
import weakref
import numpy as np


PINNED_MIN_BYTES = 1 << 20


def _allocate(cls, n, dtype):
    """
    Storage of a fresh EmbeddedFunction.  With a GPU present, vectors of a megabyte or more are page-locked (torch's
    caching host allocator; the tensor stays alive as the array's base), so that the solvers' host entry points
    (`solver(f)`, multi_boundary/scalar.py, vector.py) copy straight from / to the caller's array instead of through a
    staging buffer.
    """
    dtype = np.dtype(dtype)
    if n*dtype.itemsize >= PINNED_MIN_BYTES and dtype in (np.dtype(float), np.dtype(complex)):
        import torch
        if torch.cuda.is_available():
            tdt = torch.float64 if dtype == np.dtype(float) else torch.complex128
            return torch.empty(n, dtype=tdt, pin_memory=True).numpy().view(cls)
    return np.ndarray.__new__(cls, n, dtype)


def _sizes(ebdyc):
    gn = ebdyc.grid_phys.N
    rn = int(np.sum([int(np.prod(e.radial_shape)) for e in ebdyc]))
    return gn, rn


class EmbeddedFunction(np.ndarray):
    def __new__(cls, ebdyc, dtype=float, array=None, function=None, grid_value=None,
                radial_value_list=None, linear_data=None, zero=False):
        gn, rn = _sizes(ebdyc)
        if array is None:
            array = _allocate(cls, gn + rn, dtype)
        out = np.asarray(array).view(cls)
        out.ebdyc = weakref.ref(ebdyc)
        out._generate()
        if function is not None:
            out.define_via_function(function)
        if grid_value is not None and radial_value_list is not None:
            out.load_data(grid_value, radial_value_list)
        if grid_value is not None and radial_value_list is None:
            out.load_full_grid(grid_value)
        if linear_data is not None:
            out.load_linear_data(linear_data)
        if zero:
            out.zero()
        return out

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.ebdyc = getattr(obj, 'ebdyc', None)
        if self.ebdyc is not None and self.ebdyc() is not None and self.ndim == 1:
            gn, rn = _sizes(self.ebdyc())
            if self.size == gn + rn:
                self._generate()

    def _ebdyc_test(self):
        ebdyc = self.ebdyc() if self.ebdyc is not None else None
        if ebdyc is None:
            raise Exception('Underlying ebdyc has been deleted.')
        return ebdyc

    def _generate(self):
        ebdyc = self._ebdyc_test()
        self.n_grid = ebdyc.grid_phys.N
        self.n_data = self.size
        self.gslice = slice(0, self.n_grid)
        self.rslice = slice(self.n_grid, None)
        self.radial_slices = []
        self.radial_shapes = []
        start = self.n_grid
        for ebdy in ebdyc:
            rsh = ebdy.radial_shape
            end = start + int(np.prod(rsh))
            self.radial_slices.append(slice(start, end))
            self.radial_shapes.append(rsh)
            start = end
        self.radial_length = len(self.radial_shapes)

    def save(self):
        return {'linear_data': self.view(np.ndarray)}
    def full_save(self):
        return {'ebdyc_dict': self._ebdyc_test().save(), 'linear_data': self.view(np.ndarray)}

    def get_gdata(self):
        return np.ndarray.__getitem__(self, self.gslice).view(np.ndarray)
    def get_rdata(self):
        return np.ndarray.__getitem__(self, self.rslice).view(np.ndarray)

    def __getitem__(self, arg):
        if type(arg) == int:
            return np.ndarray.__getitem__(self, self.radial_slices[arg]).view(np.ndarray) \
                .reshape(self.radial_shapes[arg])
        elif isinstance(arg, str) and arg == 'grid':
            return np.ndarray.__getitem__(self, self.gslice).view(np.ndarray)
        return np.ndarray.__getitem__(self, arg).view(np.ndarray)

    def __setitem__(self, arg, value):
        if type(arg) == int:
            np.ndarray.__setitem__(self, self.radial_slices[arg], np.asarray(value).ravel())
        elif isinstance(arg, str) and arg == 'grid':
            np.ndarray.__setitem__(self, self.gslice, value)
        else:
            np.ndarray.__setitem__(self, arg, value)

    def __iter__(self):
        return (self[i] for i in range(self.radial_length))

    def min(self, *a, **k):
        return np.min(self.view(np.ndarray), *a, **k)
    def max(self, *a, **k):
        return np.max(self.view(np.ndarray), *a, **k)

    def load_data(self, grid_value, radial_value_list):
        gd = self.get_gdata()
        grid_value = np.asarray(grid_value)
        if grid_value.shape == gd.shape:
            gd[:] = grid_value
        else:
            gd[:] = grid_value[self._ebdyc_test().phys]
        for arg, rv in enumerate(radial_value_list):
            self[arg] = rv

    def load_full_grid(self, grid_value):
        ebdyc = self._ebdyc_test()
        rvl = ebdyc.interpolate_grid_to_radial(grid_value, order=5)
        self.load_data(grid_value, rvl)

    def load_linear_data(self, data):
        np.ndarray.__setitem__(self, slice(None, None), data)

    def define_via_function(self, f):
        ebdyc = self._ebdyc_test()
        self.get_gdata()[:] = f(ebdyc.grid_phys.x, ebdyc.grid_phys.y)
        for arg, ebdy in enumerate(ebdyc):
            self[arg] = f(ebdy.radial_x, ebdy.radial_y)

    def get_radial_value_list(self):
        return [self[arg] for arg in range(self.radial_length)]

    def get_components(self):
        gv = self.get_grid_value()
        return gv, gv*self._ebdyc_test().grid_step, self.get_radial_value_list()

    def get_grid_value(self, masked=False):
        ebdyc = self._ebdyc_test()
        arr = np.zeros(ebdyc.grid.shape, dtype=self.dtype)
        arr[ebdyc.phys] = self.get_gdata()
        if masked:
            arr = np.ma.array(arr, mask=ebdyc.ext)
        return arr

    def get_smoothed_grid_value(self):
        arr = self.get_grid_value()
        arr *= self._ebdyc_test().grid_step
        return arr

    def __str__(self):
        s = 'Physical Grid Values:\n' + str(self.get_gdata()) + \
            '\nRadial Values (first 5 embedded boundaries):\n'
        for i in range(min(self.radial_length, 5)):
            s += 'Embedded Boundary ' + str(i) + '\n' + str(self[i]) + '\n'
        return s
    __repr__ = __str__

    def copy(self, order='C'):
        out = EmbeddedFunction(self._ebdyc_test(), self.dtype)
        np.ndarray.__setitem__(out, slice(None, None), self.view(np.ndarray))
        return out
    def asarray(self):
        return np.array(self)
    def zero(self):
        np.ndarray.__setitem__(self, slice(None, None), 0.0)
    def integrate(self):
        return self._ebdyc_test().volume_integral(self)
    def gradient(self):
        return self._ebdyc_test().gradient(self)
    def get_ebdyc(self):
        return self._ebdyc_test()
    def extract_pnar(self):
        gv, _, rvl = self.get_components()
        gvp = gv[self._ebdyc_test().phys_not_in_annulus]
        return np.concatenate([gvp] + [rv.ravel() for rv in rvl])


def LoadEmbeddedFunction(d, ebdyc=None):
    from .ebdy_collection import LoadEmbeddedBoundaryCollection
    if ebdyc is None:
        if 'ebdyc_dict' not in d:
            raise Exception('Need ebdyc provided unless save was generated with full_save method.')
        ebdyc = LoadEmbeddedBoundaryCollection(d['ebdyc_dict'])
    ef = EmbeddedFunction(ebdyc)
    ef.load_linear_data(d['linear_data'])
    return ef, ebdyc


class BoundaryFunction(np.ndarray):
    """Concatenation of per-boundary value vectors (embedded_function.py:231-342)."""
    def __new__(cls, ebdyc, dtype=float, array=None, function=None, functions=None,
                data=None, zero=False):
        n = int(np.sum([ebdy.bdy.N for ebdy in ebdyc]))
        if array is None:
            array = np.ndarray.__new__(cls, n, dtype)
        out = np.asarray(array).view(cls)
        out.ebdyc = weakref.ref(ebdyc)
        out._generate()
        if function is not None:
            out.define_via_function(function)
        if functions is not None:
            out.define_via_functions(functions)
        if data is not None:
            out.load_data(data)
        if zero:
            np.ndarray.__setitem__(out, slice(None, None), 0.0)
        return out

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.ebdyc = getattr(obj, 'ebdyc', None)
        if self.ebdyc is not None and self.ebdyc() is not None and self.ndim == 1:
            if self.size == int(np.sum([e.bdy.N for e in self.ebdyc()])):
                self._generate()

    def _generate(self):
        ebdyc = self.ebdyc()
        self.slices = []
        start = 0
        for ebdy in ebdyc:
            self.slices.append(slice(start, start + ebdy.bdy.N))
            start += ebdy.bdy.N
        self.n_boundaries = len(self.slices)

    def __getitem__(self, arg):
        if type(arg) == int:
            return np.ndarray.__getitem__(self, self.slices[arg]).view(np.ndarray)
        return np.ndarray.__getitem__(self, arg).view(np.ndarray)

    def __setitem__(self, arg, value):
        if type(arg) == int:
            arg = self.slices[arg]
        np.ndarray.__setitem__(self, arg, value)

    def __iter__(self):
        return (self[i] for i in range(self.n_boundaries))

    def load_data(self, bdy_value_list):
        for i, v in enumerate(bdy_value_list):
            self[i] = v
    def define_via_function(self, f):
        for i, ebdy in enumerate(self.ebdyc()):
            self[i] = f(ebdy.bdy.x, ebdy.bdy.y)
    def define_via_functions(self, f_list):
        for i, (f, ebdy) in enumerate(zip(f_list, self.ebdyc())):
            self[i] = f(ebdy.bdy.x, ebdy.bdy.y)
    def aggregate(self):
        return np.array(self)
    def get_value_list(self):
        return [self[i] for i in range(self.n_boundaries)]
    def copy(self, order='C'):
        out = BoundaryFunction(self.ebdyc(), self.dtype)
        np.ndarray.__setitem__(out, slice(None, None), self.view(np.ndarray))
        return out
