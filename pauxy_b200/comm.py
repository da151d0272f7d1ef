"""Communicators for the host driver.

`SingleComm` plays the role of pauxy.qmc.comm.FakeComm (pauxy/qmc/comm.py:1-25)
for one process; `TorchComm` maps the collectives the hot path needs
 onto torch.distributed -- NCCL over NVLink on GPUs,
gloo in the CPU tests of the host logic.
"""
import torch


class SingleComm(object):
    rank = 0
    size = 1

    def barrier(self):
        pass

    Barrier = barrier

    def bcast(self, obj, root=0):
        return obj

    def allgather_tensor(self, t):
        return t.clone()

    def allreduce_sum_(self, t):
        return t

    def exchange(self, sends, recvs):
        assert not sends and not recvs

    def stream_barrier(self, device=None):
        pass

    def warmup(self, device=None):
        pass


class TorchComm(object):
    """One process per GPU; rank r owns global walkers [r*nw, (r+1)*nw)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)

    def barrier(self):
        self.dist.barrier(self.group)

    Barrier = barrier

    def warmup(self, device):
        """Open every channel the hot path can use (all-gather, all-reduce, send/recv with every
        peer) so that connection set-up does not land inside a timed or latency-sensitive step."""
        t = torch.zeros(8, dtype=torch.float64, device=device)
        self.allgather_tensor(t)
        self.allreduce_sum_(t)
        for shift in range(1, self.size):
            to, frm = (self.rank + shift) % self.size, (self.rank - shift) % self.size
            r = torch.empty(8, dtype=torch.float64, device=device)
            self.exchange([(to, t)], [(frm, r)])
        if device is not None and torch.device(device).type == 'cuda':
            torch.cuda.synchronize(device)

    def stream_barrier(self, device):
        """Cross-device barrier in STREAM order (no host wait): a 1-element all-reduce.  Work
        enqueued after it on any rank starts only when every rank's earlier work has finished."""
        if getattr(self, '_flag', None) is None or self._flag.device != torch.device(device):
            self._flag = torch.zeros(1, dtype=torch.float64, device=device)
        self.dist.all_reduce(self._flag, op=self.dist.ReduceOp.SUM, group=self.group)

    def bcast(self, obj, root=0):
        box = [obj]
        self.dist.broadcast_object_list(box, src=root, group=self.group)
        return box[0]

    def allgather_tensor(self, t):
        """walkers/handler.py:232 Allgather of the per-rank |weights|."""
        out = torch.empty((self.size * t.numel(),), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def allreduce_sum_(self, t):
        """estimators/mixed.py:261 Reduce + :273 bcast, as one all-reduce."""
        if t.is_complex():
            r = torch.view_as_real(t)
            self.dist.all_reduce(r, op=self.dist.ReduceOp.SUM, group=self.group)
        else:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def exchange(self, sends, recvs):
        """walkers/handler.py:301-334 Isend/Recv of walker buffers.
        sends / recvs: lists of (peer_rank, tensor).  Posted in list order on
        both sides, so messages between one pair of ranks match in order."""
        ops = []
        for peer, buf in sends:
            ops.append(self.dist.P2POp(self.dist.isend, buf, peer, group=self.group))
        for peer, buf in recvs:
            ops.append(self.dist.P2POp(self.dist.irecv, buf, peer, group=self.group))
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()
