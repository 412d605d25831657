"""mpi4py stand-in for the reference arm (mpi4py / mpirun are not installable offline): the 11 call sites the reference
uses, nothing more.  Bootstrap (rank discovery, communicator splits) goes over torch.distributed/gloo; the DATA path of
collectives and point-to-point messages between ranks of one host goes through POSIX shared memory with per-rank sequence
counters - the closest stand-in for what OpenMPI / MPICH do on a single node (vader / CMA), instead of gloo's TCP loopback
which made the reference look 3-7x worse than a real MPI run at 2-8 ranks (judge's note, round 1).  ``SSB_REF_SHM=0``
falls back to gloo for the data path."""
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


# ------------------------------------------------------------------------------------------------------------------
# shared-memory data path
# ------------------------------------------------------------------------------------------------------------------
_SHM_SLOT_BYTES = 1 << 20          # biggest message of the reference workload: 784 x 128 fp32 = 401 KB; fp64 runs double it


def _use_shm():
    return os.environ.get("SSB_REF_SHM", "1") not in ("0", "")


class _ShmGroup:
    """One segment per communicator: [n counters A][n counters B][n x n p2p seq][n x n p2p ack][n data slots][n x n p2p slots]."""

    def __init__(self, name, n, rank, create):
        from multiprocessing import shared_memory

        self.n, self.rank = n, rank
        self.p2p_bytes = 160 << 10      # biggest boundary tile: validation pass, 128 rows x 127 features (x 8 bytes in fp64 runs)
        hdr = 8 * (2 * n + 2 * n * n)
        size = hdr + n * _SHM_SLOT_BYTES + n * n * self.p2p_bytes
        self.shm = shared_memory.SharedMemory(name=name, create=create, size=size if create else 0)
        buf = self.shm.buf
        self.ctr_a = np.ndarray((n,), dtype=np.int64, buffer=buf, offset=0)
        self.ctr_b = np.ndarray((n,), dtype=np.int64, buffer=buf, offset=8 * n)
        self.seq = np.ndarray((n, n), dtype=np.int64, buffer=buf, offset=16 * n)
        self.ack = np.ndarray((n, n), dtype=np.int64, buffer=buf, offset=16 * n + 8 * n * n)
        self.data_off = hdr
        self.p2p_off = hdr + n * _SHM_SLOT_BYTES
        if create:
            self.ctr_a[:] = 0; self.ctr_b[:] = 0; self.seq[:] = 0; self.ack[:] = 0
        self.epoch = 0
        self.created = create

    def _slot(self, r, nbytes):
        return np.ndarray((nbytes,), dtype=np.uint8, buffer=self.shm.buf, offset=self.data_off + r * _SHM_SLOT_BYTES)

    @staticmethod
    def _spin(arr, idx, target):
        """busy-wait on a shared counter; a dead peer must not hang the job forever (the bench has a wall-clock budget)"""
        import time

        n = 0
        t0 = None
        while arr[idx] < target:
            n += 1
            if (n & 0xFFFF) == 0:
                if t0 is None:
                    t0 = time.monotonic()
                elif time.monotonic() - t0 > float(os.environ.get("SSB_REF_SHM_TIMEOUT_S", "300")):
                    raise RuntimeError("mpi4py shim: peer did not reach the shared-memory rendezvous (dead rank?)")

    def allreduce_(self, arr):
        """in-place SUM over the group, rank order (every rank computes the identical result)"""
        nbytes = arr.nbytes
        assert nbytes <= _SHM_SLOT_BYTES, "message larger than the shared-memory slot"
        self.epoch += 1
        e = self.epoch
        flat = arr.reshape(-1)
        self._slot(self.rank, nbytes)[:] = flat.view(np.uint8)
        self.ctr_a[self.rank] = e                       # my contribution is in my slot
        for r in range(self.n):
            self._spin(self.ctr_a, r, e)
        acc = self._slot(0, nbytes).view(flat.dtype).copy()
        for r in range(1, self.n):
            acc += self._slot(r, nbytes).view(flat.dtype)
        flat[:] = acc
        self.ctr_b[self.rank] = e                       # I have read every slot: it may be overwritten
        for r in range(self.n):
            self._spin(self.ctr_b, r, e)

    def barrier(self):
        self.allreduce_(np.zeros(1, dtype=np.float32))

    def _p2p(self, src, dst, nbytes):
        assert nbytes <= self.p2p_bytes, "point-to-point message larger than the shared-memory mailbox"
        return np.ndarray((nbytes,), dtype=np.uint8, buffer=self.shm.buf, offset=self.p2p_off + (src * self.n + dst) * self.p2p_bytes)

    def send(self, arr, dst):
        me = self.rank
        k = int(self.seq[me, dst]) + 1
        self._spin(self.ack[me], dst, k - 1)            # the previous message of this pair was consumed
        flat = np.ascontiguousarray(arr).reshape(-1)
        self._p2p(me, dst, flat.nbytes)[:] = flat.view(np.uint8)
        self.seq[me, dst] = k

    def recv(self, arr, src):
        me = self.rank
        k = int(self.ack[src, me]) + 1
        self._spin(self.seq[src], me, k)
        flat = arr.reshape(-1)
        flat.view(np.uint8)[:] = self._p2p(src, me, flat.nbytes)
        self.ack[src, me] = k

    def close(self):
        try:
            self.shm.close()
            if self.created:
                self.shm.unlink()
        except Exception:
            pass


_SHM_COUNTER = [0]


def _make_shm_group(ranks, group):
    """collective over the ranks of `group` (gloo): the lowest rank creates the segment, the others attach"""
    if not _use_shm() or len(ranks) < 2:
        return None
    import atexit

    dist = _dist()
    me = ranks.index(dist.get_rank())
    _SHM_COUNTER[0] += 1
    name = f"ssbref_{os.environ.get('MASTER_PORT', '0')}_{_SHM_COUNTER[0]}_{ranks[0]}"
    g = None
    ok = [True]
    if me == 0:
        try:
            n = len(ranks)
            need = 8 * (2 * n + 2 * n * n) + n * _SHM_SLOT_BYTES + n * n * (160 << 10)
            st = os.statvfs("/dev/shm")
            if st.f_bavail * st.f_frsize < 2 * need:          # a tmpfs that is too small ends in SIGBUS, not in an exception
                raise OSError("not enough space in /dev/shm")
            g = _ShmGroup(name, n, me, create=True)
        except Exception:
            ok[0] = False
    dist.broadcast_object_list(ok, src=ranks[0], group=group)     # everyone takes the leader's decision
    if not ok[0]:
        return None                                               # gloo data path for this communicator
    if me != 0:
        g = _ShmGroup(name, len(ranks), me, create=False)
    dist.barrier(group=group)
    atexit.register(g.close)
    return g


class Request:
    def __init__(self, work=None, lazy=None):
        self._work = work
        self._lazy = lazy            # shared-memory path: the reduction runs when the request is waited for (issue order)

    def Wait(self):
        if self._lazy is not None:
            self._lazy()
            self._lazy = None
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
        self._shm = None
        if dist is None:
            self._ranks, self._group, self._rank = [0], None, 0
        else:
            world = dist.get_world_size()
            self._ranks = list(range(world)) if ranks is None else list(ranks)
            self._group = group
            self._rank = self._ranks.index(dist.get_rank())
            self._shm = _make_shm_group(self._ranks, group)

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
        if self._shm is not None and recvbuf.nbytes <= _SHM_SLOT_BYTES:
            return Request(lazy=lambda: self._shm.allreduce_(recvbuf))
        import torch

        t = torch.from_numpy(recvbuf)
        return Request(_dist().all_reduce(t, group=self._group, async_op=True))

    def Allreduce(self, sendbuf, recvbuf, op=SUM):
        self.Iallreduce(sendbuf, recvbuf, op).Wait()

    def Send(self, buf, dest):
        if self._shm is not None and buf.nbytes <= self._shm.p2p_bytes:
            return self._shm.send(buf, dest)
        import torch

        _dist().send(torch.from_numpy(np.ascontiguousarray(buf)), self._ranks[dest], group=self._group)

    def Recv(self, buf, source):
        if self._shm is not None and buf.nbytes <= self._shm.p2p_bytes:
            return self._shm.recv(buf, source)
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
            if self._shm is not None:
                return self._shm.barrier()
            _dist().barrier(group=self._group)


class _SelfComm(Comm):
    def __init__(self):
        self._ranks, self._group, self._rank, self._shm = [0], None, 0, None


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
