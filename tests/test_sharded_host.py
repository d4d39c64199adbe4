"""Host-side plumbing of the sharded run (polychordlite_b200/mgpu.py) on CPU: two gloo ranks exchange the
IPC handles in rank order, call the engine with identical settings and sum the per-rank evaluation counts.
The CUDA calls are replaced by a recording stand-in (there is no GPU here); the device side is covered by
tests/test_gpu_sharded.py on a 2-GPU box."""
import os
import socket

import torch.distributed as dist
import torch.multiprocessing as mp

from polychordlite_b200 import mgpu


def test_chain_dealing_covers_every_chain_once():
    for K, world in [(250, 2), (250, 8), (7, 4), (3, 8)]:
        seen = sorted(k for r in range(world) for k in mgpu.chains_of_rank(K, r, world))
        assert seen == list(range(K))
        assert all(k % world == r for r in range(world) for k in mgpu.chains_of_rank(K, r, world))


class FakeCapi:
    def __init__(self, rank):
        self.rank, self.log = rank, []

    def mgpu_create(self, settings, world):
        self.log.append(("create", world))
        return bytes([self.rank]) * 64

    def mgpu_attach(self, rank, world, handles):
        self.log.append(("attach", rank, world, [h[0] for h in handles]))

    def mgpu_destroy(self):
        self.log.append(("destroy",))

    def run(self, settings, **kw):
        self.log.append(("run", settings))

        class Info:
            nlike = 1000 + self.rank
        return Info(), []


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    capi = FakeCapi(rank)
    info, dumps, total = mgpu.run_sharded("settings", capi=capi)
    q.put((rank, capi.log, info.nlike, total))
    dist.destroy_process_group()


def test_two_ranks_exchange_handles_and_sum_counts():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank, log, nlike, total in out:
        assert log[0] == ("create", 2)
        assert log[1] == ("attach", rank, 2, [0, 1])       # handles arrive in rank order
        assert log[2] == ("run", "settings") and log[3] == ("destroy",)
        assert nlike == 1000 + rank and total == 2001
