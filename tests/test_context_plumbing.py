"""Host-side orchestration of the symmetric-memory contexts (NVLS multicast object, peer-memory pipeline slots),
exercised on CPU with gloo process groups and fake native classes: who creates / imports what, which neighbour's
handles are opened, how the file descriptor travels."""
import os

import pytest
import torch
import torch.multiprocessing as mp


class _FakeModule:
    """Stands in for shallowspeed_b200._C inside the spawned ranks."""

    log = None

    class PpContext:
        def __init__(self, n_mu, mb, ld_in, ld_out, first, last):
            self.args = (n_mu, mb, ld_in, ld_out, first, last)
            self.prev = self.next = None

        def export_handles(self):
            return f"handles-of-{os.environ['RANK']}".encode()

        def open_prev(self, h):
            self.prev = bytes(h)

        def open_next(self, h):
            self.next = bytes(h)

    class NvlsContext:
        @staticmethod
        def supported():
            return True

        def __init__(self, dp, rank, numel, lr):
            self.dp, self.rank, self.numel = dp, rank, numel
            self.calls = []
            self.w, self.g = torch.zeros(numel), torch.zeros(numel)

        def export_fd(self):
            self.calls.append("export")
            r, w = os.pipe()
            os.write(w, b"multicast-object")
            os.close(w)
            return r

        def import_fd(self, fd):
            self.calls.append(("import", os.read(fd, 64)))

        def add_device(self):
            self.calls.append("add")

        def bind_and_map(self):
            self.calls.append("bind")

        def weights(self):
            return self.w

        def grads(self):
            return self.g


class _Eng:
    def __init__(self, lds):
        self._lds = lds

    def boundary_lds(self):
        return self._lds


class _Arena:
    def __init__(self, n):
        self.weights, self.grads, self.numel = torch.arange(n, dtype=torch.float32), torch.ones(n), n

    def rebind(self, w, g, copy=True):
        if copy:
            w[: self.numel].copy_(self.weights)
            g[: self.numel].copy_(self.grads)
        self.weights, self.grads = w, g


class _Model:
    def __init__(self, n):
        self.arena = _Arena(n)


def _rank_main(rank, world, port, out_dir, what):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from shallowspeed_b200.parallel import engine as E
    from shallowspeed_b200.parallel.comm import ProcessGrid, make_torch_comms

    dist.init_process_group("gloo", rank=rank, world_size=world)
    E._C = lambda: _FakeModule
    if what == "pp":
        _dp, pp_comm = make_torch_comms(ProcessGrid(1, world, rank))
        lds = [(784, 128), (128, 128), (128, 16)][rank]
        ctx = E.make_pp_context(pp_comm, _Eng(lds), 4, 32, rank == 0, rank == world - 1)
        torch.save({"args": ctx.args, "prev": ctx.prev, "next": ctx.next}, os.path.join(out_dir, f"r{rank}.pt"))
    else:
        dp_comm, _pp = make_torch_comms(ProcessGrid(world, 1, rank))
        model = _Model(10)
        ctx = E.make_nvls_context(dp_comm, model, 0.1)
        torch.save({"calls": ctx.calls, "w": model.arena.weights.clone(), "same": model.arena.weights is ctx.w},
                   os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _spawn(world, what, tmp_path):
    port = 29900 + (os.getpid() % 90) + (7 if what == "pp" else 0)
    mp.spawn(_rank_main, args=(world, port, str(tmp_path), what), nprocs=world, join=True)
    return [torch.load(tmp_path / f"r{r}.pt", weights_only=False) for r in range(world)]


def test_pp_context_opens_the_right_neighbours(tmp_path):
    r = _spawn(3, "pp", tmp_path)
    assert r[0]["prev"] is None and r[0]["next"] == b"handles-of-1"
    assert r[1]["prev"] == b"handles-of-0" and r[1]["next"] == b"handles-of-2"
    assert r[2]["prev"] == b"handles-of-1" and r[2]["next"] is None
    assert r[1]["args"] == (4, 32, 128, 128, False, False)


def test_nvls_context_leader_exports_peers_import_the_same_object(tmp_path):
    r = _spawn(2, "nvls", tmp_path)
    assert r[0]["calls"] == ["export", "add", "bind"]
    assert r[1]["calls"] == [("import", b"multicast-object"), "add", "bind"]
    for x in r:
        assert x["same"] and torch.equal(x["w"], torch.arange(10, dtype=torch.float32))   # arena moved, values kept
