"""world_size-2 gloo test (CPU) of the N>1 host path: round-robin sharding of triples, the one-time
parameter broadcast, max-over-ranks timing and the counter gather.  The step itself has no collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hairfastgan_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import hairfastgan_b200.model as M
        torch.manual_seed(1234 + rank)                      # ranks start with DIFFERENT weights
        gen = M.Generator(64, 512, 2)
        before = gen.conv1.conv.weight.detach().clone()
        nbytes = sharding.broadcast_module_(gen, src=0)
        mine = sharding.shard_indices(7, rank, world)
        tmax = sharding.max_over_ranks(1.0 + rank)
        counts = sharding.gather_counts(len(mine))
        ret[rank] = dict(w=gen.conv1.conv.weight.detach().clone(), before=before, nbytes=nbytes, mine=mine,
                         tmax=tmax, counts=counts,
                         noise=gen.noises.noise_3.clone())
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_broadcast():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    r0, r1 = ret[0], ret[1]
    assert not torch.equal(r0["before"], r1["before"])
    assert torch.equal(r0["w"], r1["w"]) and torch.equal(r0["w"], r0["before"])     # rank 0's weights win
    assert torch.equal(r0["noise"], r1["noise"])                                     # buffers too
    assert r0["nbytes"] == r1["nbytes"] > 0
    assert sorted(r0["mine"] + r1["mine"]) == list(range(7)) and not set(r0["mine"]) & set(r1["mine"])
    assert r0["tmax"] == r1["tmax"] == 2.0
    assert list(r0["counts"]) == [4, 3] == list(r1["counts"])


def test_single_process_paths_are_noops():
    assert sharding.shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    assert sharding.max_over_ranks(3.5) == 3.5
    assert sharding.gather_counts(3) == [3]
