"""Test-infrastructure stub for mpi4py (absent from this image): one rank.

Used only by oracle/ref_harness.py to run the unmodified reference.
"""
from . import MPI  # noqa: F401


class _RC(object):
    """`mpi4py.rc.recv_mprobe = False` is set at import by
    /root/reference/pauxy/estimators/back_propagation.py:5."""
    recv_mprobe = False


rc = _RC()
