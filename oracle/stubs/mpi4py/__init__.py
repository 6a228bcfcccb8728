"""Single-rank stand-in for mpi4py, used ONLY to import the read-only reference
(/root/reference) in the build container so it can act as the parity oracle.
Test infrastructure -- never imported by pyhype_b200."""
from . import MPI  # noqa: F401
