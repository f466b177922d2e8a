"""
ctypes binding of libipde_b200.so (the C ABI declared in include/ipde_b200.h).

The prototypes are parsed from the header itself, so the header is the single source of
truth.  There is no CPU fallback: if the library is missing, importing this module raises.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libipde_b200.so')
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'ipde_b200.h')

_SCALARS = {'int': ctypes.c_int, 'double': ctypes.c_double, 'int64_t': ctypes.c_int64}


def parse_header(path=HEADER_PATH):
    """Return {name: (restype, [argtypes])} for every function the header declares."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    src = re.sub(r'//[^\n]*', ' ', src)
    protos = {}
    for m in re.finditer(r'\b(const\s+char\s*\*|int64_t|int)\s+(ipde_\w+)\s*\(([^)]*)\)\s*;', src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if 'char' in ret else (ctypes.c_int64 if ret == 'int64_t' else ctypes.c_int)
        argtypes = []
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                if '*' in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    base = a.replace('const', '').split()[0]
                    argtypes.append(_SCALARS[base])
        protos[name] = (restype, argtypes)
    return protos


class IpdeError(Exception):
    pass


class _Lib(object):
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                'ipde_b200: %s not found -- build it with `python -c "import __graft_entry__ as g; g.build()"` '
                '(there is no CPU fallback)' % LIB_PATH)
        self._dll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(self._dll, name)       # raises AttributeError if the symbol is missing
            fn.restype = restype
            fn.argtypes = argtypes

    def last_error(self):
        return self._dll.ipde_last_error().decode()

    def __getattr__(self, name):
        if not name.startswith('ipde_'):
            raise AttributeError(name)
        fn = getattr(self._dll, name)
        restype, _ = self.protos[name]
        if restype is not ctypes.c_int or name in ('ipde_version', 'ipde_fft_next_size', 'ipde_fft_size_supported',
                                                   'ipde_device_sm_count'):
            return fn

        def checked(*args):
            rc = fn(*args)
            if rc != 0:
                raise IpdeError('%s failed (%d): %s' % (name, rc, self.last_error()))
            return rc
        self.__dict__[name] = checked
        return checked


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
