import os

import numpy as np

IN_PLACE = "IN_PLACE"
SUM = "SUM"


def _dist():
    import torch.distributed as dist

    return dist


def _ensure_init():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return None
    dist = _dist()
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=world)
    return dist


class Request:
    def __init__(self, work=None):
        self._work = work

    def Wait(self):
        if self._work is not None:
            self._work.wait()
            self._work = None

    @staticmethod
    def Waitall(requests):
        for r in requests:
            r.Wait()


class Comm:
    def __init__(self, ranks=None, group=None):
        dist = _ensure_init()
        if dist is None:
            self._ranks, self._group, self._rank = [0], None, 0
        else:
            world = dist.get_world_size()
            self._ranks = list(range(world)) if ranks is None else list(ranks)
            self._group = group
            self._rank = self._ranks.index(dist.get_rank())

    # -- introspection
    @property
    def size(self):
        return len(self._ranks)

    @property
    def rank(self):
        return self._rank

    def Get_rank(self):
        return self._rank

    def Get_size(self):
        return len(self._ranks)

    # -- communicator management
    def Split(self, color=0, key=0):
        if self.size == 1:
            return _SelfComm()
        dist = _dist()
        colors = [None] * self.size
        dist.all_gather_object(colors, int(color), group=self._group)
        mine = None
        for c in sorted(set(colors)):
            ranks = [self._ranks[i] for i, ci in enumerate(colors) if ci == c]
            g = dist.new_group(ranks, backend="gloo")     # collective over the world: all ranks create all groups
            if c == int(color):
                mine = Comm(ranks, g)
        return mine

    # -- collectives / p2p on numpy buffers
    def Iallreduce(self, sendbuf, recvbuf, op=SUM):
        assert sendbuf is IN_PLACE and op is SUM
        if self.size == 1:
            return Request()
        import torch

        t = torch.from_numpy(recvbuf)
        return Request(_dist().all_reduce(t, group=self._group, async_op=True))

    def Allreduce(self, sendbuf, recvbuf, op=SUM):
        self.Iallreduce(sendbuf, recvbuf, op).Wait()

    def Send(self, buf, dest):
        import torch

        _dist().send(torch.from_numpy(np.ascontiguousarray(buf)), self._ranks[dest], group=self._group)

    def Recv(self, buf, source):
        import torch

        _dist().recv(torch.from_numpy(buf), self._ranks[source], group=self._group)

    def gather(self, obj, root=0):
        if self.size == 1:
            return [obj]
        out = [None] * self.size
        _dist().all_gather_object(out, obj, group=self._group)
        return out if self._rank == root else None

    def Barrier(self):
        if self.size > 1:
            _dist().barrier(group=self._group)


class _SelfComm(Comm):
    def __init__(self):
        self._ranks, self._group, self._rank = [0], None, 0


class _World:
    """Lazily initialised COMM_WORLD."""

    _comm = None

    def _get(self):
        if _World._comm is None:
            _World._comm = Comm()
        return _World._comm

    def __getattr__(self, name):
        return getattr(self._get(), name)


COMM_WORLD = _World()
