"""CPU, world_size 2 over gloo: the batch-sharding host logic of the multi-GPU path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pvnet_b200 import distributed as pd


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 16, 128, 129):
        for world in (1, 2, 4, 8):
            spans = [pd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_image_seed_independent_of_world_size():
    n = 32
    ref = [pd.image_seed(5, i) for i in range(n)]
    for world in (2, 4, 8):
        got = []
        for r in range(world):
            lo, hi = pd.shard_range(n, r, world)
            got += [pd.image_seed(5, i) for i in range(lo, hi)]
        assert got == ref
    assert len(set(ref)) == n


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = pd.shard_range(n_items, rank, world)
    full = torch.arange(n_items * 9 * 2, dtype=torch.float32).view(n_items, 9, 2)
    out = pd.gather_results(full[lo:hi].clone(), n_items)
    out2 = pd.gather_results(full[lo:hi].clone())          # sizes discovered by a first all_gather
    q.put((rank, torch.equal(out, full) and torch.equal(out2, full)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [16, 7])
def test_gather_results_world2_gloo(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_gather_results_single_process_is_identity():
    t = torch.randn(4, 9, 2)
    assert pd.gather_results(t) is t
