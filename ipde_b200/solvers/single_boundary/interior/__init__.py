from .modified_helmholtz import ModifiedHelmholtzSolver
from .poisson import PoissonSolver
from .stokes import StokesSolver
