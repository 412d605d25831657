"""Minimal mpi4py-compatible shim (NOT product code).

mpi4py / mpirun cannot be installed offline in this environment, so
the UNMODIFIED reference (installed under baseline/_ref) is driven through this shim: it
implements exactly the mpi4py surface the reference touches on
top of torch.distributed's gloo backend (host memory, like MPI).  World size 1 needs no
process group at all."""
from . import MPI  # noqa: F401
