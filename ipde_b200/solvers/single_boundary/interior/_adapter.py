"""Shared plumbing of the single-boundary adapters: a one-boundary collection over the boundary's grid."""
import numpy as np
from ....ebdy_collection import EmbeddedBoundaryCollection
from ....embedded_function import EmbeddedFunction


def collection_of(ebdy):
    """
    The older single-boundary solvers of the reference take an EmbeddedBoundary that was registered on a grid by
    itself (`ebdy.register_grid(grid)`, ipde/embedded_boundary.py:185).  Here registration is a collection service,
    so the adapter accepts (a) a registered 1-boundary EmbeddedBoundaryCollection, (b) an EmbeddedBoundary that
    belongs to one (`ebdy.ebdyc`), or (c) an EmbeddedBoundary carrying `ebdy.grid` -- for which a collection is
    built and registered.
    """
    if isinstance(ebdy, EmbeddedBoundaryCollection):
        if ebdy.N != 1:
            raise Exception('single-boundary solver needs a collection with exactly one boundary')
        return ebdy
    owner = getattr(ebdy, 'ebdyc', None)
    if owner is not None:
        if owner.N != 1:
            raise Exception('this EmbeddedBoundary belongs to a multi-boundary collection: use the multi_boundary solvers')
        return owner
    grid = getattr(ebdy, 'grid', None)
    if grid is None:
        raise Exception('EmbeddedBoundary has no grid: register it (EmbeddedBoundaryCollection([ebdy]).register_grid(grid))')
    ebdyc = EmbeddedBoundaryCollection([ebdy])
    ebdyc.register_grid(grid)
    ebdy.ebdyc = ebdyc
    return ebdyc


def pack(ebdyc, f, fr):
    return EmbeddedFunction(ebdyc, grid_value=np.asarray(f, dtype=float), radial_value_list=[np.asarray(fr, dtype=float)])


def unpack(u):
    return u.get_grid_value(), np.array(u[0])
