"""Single-rank loop-back communicator.  Isend/Recv loop back by tag because
the reference's comb clones walkers through Isend/Recv even on one rank
(/root/reference/pauxy/walkers/handler.py:301-325).  No `Win` attribute on
purpose: the reference then falls back to numpy.zeros
(/root/reference/pauxy/utils/mpi.py:32-35)."""
import copy
import numpy

SUM = 'sum'
COMM_TYPE_SHARED = 0


class _Request(object):
    def wait(self):
        return None

    Wait = wait


class _Comm(object):
    rank = 0
    size = 1

    def __init__(self):
        self._queue = {}

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def Split_type(self, *a, **k):
        return self

    def barrier(self):
        pass

    Barrier = barrier

    def bcast(self, obj, root=0):
        return obj

    def Bcast(self, buf, root=0):
        return None

    def gather(self, obj, root=0):
        return [obj]

    def scatter(self, obj, root=0):
        return obj[0]

    def Allgather(self, send, recv):
        numpy.copyto(numpy.asarray(recv).reshape(numpy.asarray(send).shape), send)

    def Allreduce(self, send, recv, op=None):
        numpy.copyto(recv, send)

    def Reduce(self, send, recv, op=None, root=0):
        numpy.copyto(recv, send)

    def Isend(self, buf, dest=0, tag=0):
        self._queue.setdefault(int(round(float(numpy.real(tag)))), []).append(numpy.array(buf, copy=True))
        return _Request()

    def Recv(self, buf, source=0, tag=0):
        data = self._queue[int(round(float(numpy.real(tag))))].pop(0)
        numpy.copyto(buf, data)


COMM_WORLD = _Comm()
